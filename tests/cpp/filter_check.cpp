// Host-only check of the exponential filter's level schedule (natrium_b200/csrc/filter_build.h): on random continuous-element
// meshes (nx x ny [x nz] cells of (p+1)^dim DoFs sharing their faces) in lexicographic, reversed, Morton-like and shuffled cell
// order, (1) cells of one level share no DoF, (2) of two cells that share a DoF the earlier one sits in a lower level, and
// (3) applying a non-commuting cell update level by level (cells of a level in any order) gives bit for bit what the sequential
// loop over the cells gives -- the property the device relies on.
// Usage: filter_check <dim> <nx> <ny> <nz> <p> <order 0=lex 1=reversed 2=morton 3=shuffled> <seed>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <random>
#include "../../natrium_b200/csrc/filter_build.h"

static uint32_t morton(uint32_t x, uint32_t y, uint32_t z)
{
    uint32_t m = 0;
    for (int b = 0; b < 10; b++) m |= ((x >> b) & 1u) << (3 * b) | ((y >> b) & 1u) << (3 * b + 1) | ((z >> b) & 1u) << (3 * b + 2);
    return m;
}

int main(int argc, char** argv)
{
    const int dim = argc > 1 ? atoi(argv[1]) : 2;
    int nc[3] = {argc > 2 ? atoi(argv[2]) : 4, argc > 3 ? atoi(argv[3]) : 3, argc > 4 ? atoi(argv[4]) : 2};
    const int p = argc > 5 ? atoi(argv[5]) : 2, order = argc > 6 ? atoi(argv[6]) : 0;
    std::mt19937 rng(argc > 7 ? (unsigned)atoi(argv[7]) : 1u);
    if (dim == 2) nc[2] = 1;
    int nd[3];
    for (int j = 0; j < 3; j++) nd[j] = j < dim ? nc[j] * p + 1 : 1;
    const int64_t nloc = (int64_t)nd[0] * nd[1] * nd[2];
    const int n1 = p + 1, n = dim == 2 ? n1 * n1 : n1 * n1 * n1;
    const int64_t n_cells = (int64_t)nc[0] * nc[1] * nc[2];
    std::vector<int64_t> cells((size_t)n_cells);
    std::iota(cells.begin(), cells.end(), 0);
    auto coords = [&](int64_t c, int* o) { o[0] = (int)(c % nc[0]); o[1] = (int)((c / nc[0]) % nc[1]); o[2] = (int)(c / ((int64_t)nc[0] * nc[1])); };
    if (order == 1) std::reverse(cells.begin(), cells.end());
    if (order == 2) std::sort(cells.begin(), cells.end(), [&](int64_t a, int64_t b) { int ca[3], cb[3]; coords(a, ca); coords(b, cb); return morton(ca[0], ca[1], ca[2]) < morton(cb[0], cb[1], cb[2]); });
    if (order == 3) std::shuffle(cells.begin(), cells.end(), rng);
    std::vector<int32_t> dofs((size_t)n_cells * n);
    for (int64_t k = 0; k < n_cells; k++) {
        int c[3];
        coords(cells[(size_t)k], c);
        int i = 0;
        for (int lz = 0; lz < (dim == 3 ? n1 : 1); lz++) for (int ly = 0; ly < n1; ly++) for (int lx = 0; lx < n1; lx++)
            dofs[(size_t)k * n + i++] = (int32_t)((((int64_t)(c[2] * p + lz)) * nd[1] + c[1] * p + ly) * nd[0] + c[0] * p + lx);
    }
    nbfilter::Levels L;
    if (nbfilter::build_levels(n_cells, n, dofs.data(), nloc, L) != 0) { printf("FAIL duplicate DoF reported\n"); return 1; }
    const int64_t n_levels = (int64_t)L.level_off.size() - 1;
    std::vector<int32_t> level_of((size_t)n_cells, -1);
    for (int64_t l = 0; l < n_levels; l++)
        for (int64_t i = L.level_off[(size_t)l]; i < L.level_off[(size_t)l + 1]; i++) {
            if (level_of[(size_t)L.cells[(size_t)i]] >= 0) { printf("FAIL cell in two levels\n"); return 1; }
            level_of[(size_t)L.cells[(size_t)i]] = (int32_t)l;
        }
    // (1) + (2): per DoF, the cells that contain it, in order of appearance, have strictly increasing levels
    std::vector<int32_t> last((size_t)nloc, -1);
    for (int64_t k = 0; k < n_cells; k++)
        for (int i = 0; i < n; i++) {
            int32_t& ll = last[(size_t)dofs[(size_t)k * n + i]];
            if (level_of[(size_t)k] < 0 || level_of[(size_t)k] <= ll) { printf("FAIL order of cells sharing DoF %d\n", dofs[(size_t)k * n + i]); return 1; }
            ll = level_of[(size_t)k];
        }
    // (3) a non-commuting update: v[dofs] <- M v[dofs] + cell index, M a fixed dense matrix
    std::vector<double> M((size_t)n * n);
    std::uniform_real_distribution<double> U(-0.3, 0.3);
    for (auto& m : M) m = U(rng);
    std::vector<double> v0((size_t)nloc), a, b;
    for (auto& x : v0) x = U(rng);
    auto apply = [&](std::vector<double>& v, int64_t k) {
        std::vector<double> s((size_t)n), t((size_t)n);
        for (int i = 0; i < n; i++) s[(size_t)i] = v[(size_t)dofs[(size_t)k * n + i]];
        for (int i = 0; i < n; i++) { double acc = 1e-3 * (double)k; for (int j = 0; j < n; j++) acc += M[(size_t)i * n + j] * s[(size_t)j]; t[(size_t)i] = acc; }
        for (int i = 0; i < n; i++) v[(size_t)dofs[(size_t)k * n + i]] = t[(size_t)i];
    };
    a = v0;
    for (int64_t k = 0; k < n_cells; k++) apply(a, k);
    b = v0;
    for (int64_t l = 0; l < n_levels; l++) {
        std::vector<int32_t> lv(L.cells.begin() + L.level_off[(size_t)l], L.cells.begin() + L.level_off[(size_t)l + 1]);
        std::shuffle(lv.begin(), lv.end(), rng);          // any order inside a level
        for (int32_t k : lv) apply(b, k);
    }
    for (int64_t i = 0; i < nloc; i++) if (a[(size_t)i] != b[(size_t)i]) { printf("FAIL level-wise result differs at DoF %lld\n", (long long)i); return 1; }
    // a cell that lists a DoF twice is refused
    std::vector<int32_t> bad(dofs.begin(), dofs.begin() + n);
    bad[1] = bad[0];
    nbfilter::Levels L2;
    if (nbfilter::build_levels(1, n, bad.data(), nloc, L2) != 1) { printf("FAIL duplicate DoF not reported\n"); return 1; }
    printf("OK cells=%lld levels=%lld largest_level=%lld\n", (long long)n_cells, (long long)n_levels,
           (long long)[&] { int64_t m = 0; for (int64_t l = 0; l < n_levels; l++) m = std::max(m, L.level_off[(size_t)l + 1] - L.level_off[(size_t)l]); return m; }());
    return 0;
}
