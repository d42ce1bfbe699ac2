// Host-only check of the NB_FMT_GRID driving tables (natrium_b200/csrc/grid_build.h): assembles a semi-Lagrangian-like
// matrix on a periodic tensor grid (continuous FE_Q(p): cells share their face points, periodic faces are distinct points),
// with an arbitrary DoF numbering and shuffled row entries, runs the same steps as nb200_upload_block_csr /
// nb200_finalize_matrix (row sort by grid position -> dictionary -> grid tables) and replays the kernel's access pattern on
// the CPU: TMA boxes (out-of-range points read 0) -> staging buffer, row = sum_k W[pattern][k] * xs[offset + off_table[k]],
// rows marked generic from their dictionary list.  Compares with the plain CSR product.
// Usage: grid_check <dim> <cells_x> <cells_y> <cells_z> <p> <numbering 0=lex 1=random> <cap> <seed> [wall]
//        wall = 1 adds a few off-diagonal (bounce-back like) entries and truncated rows, which must come out generic.
//        grid_check file <path> <cap>: the problem comes from a dump (tests/test_harness_and_halo.py: any rank of a partitioned
//        harness problem, ghost slots included): int64 header {dim, p, n_owned, n_ghost, ndir, dims[3]}, int32 coords[nloc][dim],
//        then per direction int64 nnz, int64 rowptr[n_owned + 1], int32 col[nnz], double val[nnz].
#include <cstdio>
#include <cstdlib>
#include <array>
#include <numeric>
#include <random>
#include <string>
#include <cmath>
#include "../../natrium_b200/csrc/grid_build.h"

struct ExtraBlock {            // an off-diagonal block of a direction's block row (bounce-back walls): reads population bj + 1
    int bj = 0;
    std::vector<int64_t> rowptr;
    std::vector<int32_t> col;
    std::vector<double> val;
};

struct Problem {
    int dim = 3, p = 4, ndir = 0, wall = 0;
    int64_t n = 0, nloc = 0, stride = 0;          // rows (owned), owned + ghost, population pitch
    nbgrid::Grid g;
    std::vector<double> x;
    std::vector<nbdict::DirBuild> dirs;
    std::vector<std::vector<int64_t>> rowptr, wall_ptr;
    std::vector<std::vector<int32_t>> col, wall_col;
    std::vector<std::vector<double>> val, wall_val;
    std::vector<std::vector<ExtraBlock>> extra;       // file mode, multi-block format
};

static bool add_direction(Problem& P, int a)
{
    // what nb200_upload_block_csr does: sort by grid position, then the dictionary
    std::vector<int32_t> sc;
    std::vector<double> sv;
    const int32_t* cp = P.col[(size_t)a].data();
    const double* vp = P.val[(size_t)a].data();
    if (nbgrid::sort_rows_by_grid(P.g.gidx_of_int, P.n, P.rowptr[(size_t)a].data(), cp, vp, sc, sv)) { cp = sc.data(); vp = sv.data(); }
    const char* msg = "";
    if (!nbdict::add_block(P.dirs[(size_t)a], P.n, P.rowptr[(size_t)a].data(), cp, vp, (int64_t)(a + 1) * P.stride, getenv("GRID_CHECK_TOL") ? atof(getenv("GRID_CHECK_TOL")) : 0.0, 63, (1 << 26) - 1, &msg)) { printf("FAIL add_block %s\n", msg); return false; }
    if (P.wall) {
        const int b = (a + 1) % P.ndir;
        if (!nbdict::add_block(P.dirs[(size_t)a], P.n, P.wall_ptr[(size_t)a].data(), P.wall_col[(size_t)a].data(), P.wall_val[(size_t)a].data(),
                               (int64_t)(b + 1) * P.stride, 0.0, 63, (1 << 26) - 1, &msg)) { printf("FAIL add_block wall %s\n", msg); return false; }
    }
    return true;
}

static bool load_file(const char* path, Problem& P)
{
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    int64_t h[8];
    if (fread(h, 8, 8, f) != 8) return false;
    const bool multi = h[4] < 0;        // negative direction count: every direction lists its blocks {nblk; per block: bj, nnz, rowptr, col, val}
    P.dim = (int)h[0]; P.p = (int)h[1]; P.n = h[2]; P.nloc = h[2] + h[3]; P.ndir = (int)(multi ? -h[4] : h[4]);
    P.stride = ((P.nloc + 31) / 32) * 32;
    nbgrid::Grid& g = P.g;
    g.dim = P.dim; g.fe_order = P.p;
    g.xshift = 1;                                   // as nb200_set_dof_grid does when the cells are declared (fe_order > 0)
    for (int j = 0; j < 3; j++) g.n[j] = j < P.dim ? (int32_t)h[5 + j] + (j == 0 ? g.xshift : 0) : 1;
    g.nxp = (g.n[0] + 1) & ~1;
    g.G = g.nxp * g.n[1] * g.n[2];
    std::vector<int32_t> coords((size_t)P.nloc * P.dim);
    if (fread(coords.data(), 4, coords.size(), f) != coords.size()) return false;
    g.gidx_of_int.assign((size_t)P.nloc, -1);
    for (int64_t u = 0; u < P.nloc; u++) {
        int c[3] = {0, 0, 0};
        for (int j = 0; j < P.dim; j++) c[j] = coords[(size_t)(u * P.dim + j)];
        g.gidx_of_int[(size_t)u] = (int32_t)g.flat(c[0] + g.xshift, c[1], c[2]);
    }
    P.dirs.resize((size_t)P.ndir); P.rowptr.resize((size_t)P.ndir); P.col.resize((size_t)P.ndir); P.val.resize((size_t)P.ndir);
    P.wall_ptr.resize((size_t)P.ndir); P.wall_col.resize((size_t)P.ndir); P.wall_val.resize((size_t)P.ndir);
    P.extra.resize((size_t)P.ndir);
    for (int a = 0; a < P.ndir; a++) {
        int64_t nblk = 1, bj = a, nnz;
        if (multi && fread(&nblk, 8, 1, f) != 1) return false;
        if (multi) {
            // the diagonal block first (it fills rowptr / col / val), the others go to `extra`
            std::vector<ExtraBlock> blks((size_t)nblk);
            for (auto& B : blks) {
                if (fread(&bj, 8, 1, f) != 1 || fread(&nnz, 8, 1, f) != 1) return false;
                B.bj = (int)bj;
                B.rowptr.resize((size_t)P.n + 1); B.col.resize((size_t)nnz); B.val.resize((size_t)nnz);
                if (fread(B.rowptr.data(), 8, (size_t)P.n + 1, f) != (size_t)P.n + 1) return false;
                if (nnz && (fread(B.col.data(), 4, (size_t)nnz, f) != (size_t)nnz || fread(B.val.data(), 8, (size_t)nnz, f) != (size_t)nnz)) return false;
            }
            P.dirs[(size_t)a].init(P.n);
            bool have_diag = false;
            for (auto& B : blks) {
                if (B.bj == a) { P.rowptr[(size_t)a] = B.rowptr; P.col[(size_t)a] = B.col; P.val[(size_t)a] = B.val; have_diag = true; }
            }
            if (!have_diag) { P.rowptr[(size_t)a].assign((size_t)P.n + 1, 0); }
            if (!add_direction(P, a)) return false;
            for (auto& B : blks) {
                if (B.bj == a) continue;
                std::vector<int32_t> sc; std::vector<double> sv;
                const int32_t* cp = B.col.data(); const double* vp = B.val.data();
                if (nbgrid::sort_rows_by_grid(P.g.gidx_of_int, P.n, B.rowptr.data(), cp, vp, sc, sv)) { cp = sc.data(); vp = sv.data(); }
                const char* msg = "";
                if (!nbdict::add_block(P.dirs[(size_t)a], P.n, B.rowptr.data(), cp, vp, (int64_t)(B.bj + 1) * P.stride, 0.0, 63, (1 << 26) - 1, &msg)) { printf("FAIL add_block %s\n", msg); return false; }
                P.extra[(size_t)a].push_back(B);
            }
            continue;
        }
        if (fread(&nnz, 8, 1, f) != 1) return false;
        P.rowptr[(size_t)a].resize((size_t)P.n + 1); P.col[(size_t)a].resize((size_t)nnz); P.val[(size_t)a].resize((size_t)nnz);
        if (fread(P.rowptr[(size_t)a].data(), 8, (size_t)P.n + 1, f) != (size_t)P.n + 1) return false;
        if (fread(P.col[(size_t)a].data(), 4, (size_t)nnz, f) != (size_t)nnz) return false;
        if (fread(P.val[(size_t)a].data(), 8, (size_t)nnz, f) != (size_t)nnz) return false;
        P.dirs[(size_t)a].init(P.n);
        if (!add_direction(P, a)) return false;
    }
    fclose(f);
    std::mt19937_64 rng(7);
    std::uniform_real_distribution<double> U(-1.0, 1.0);
    P.x.resize((size_t)(P.ndir + 1) * P.stride);
    for (auto& v : P.x) v = U(rng);
    return true;
}

static int check(Problem& P, int cap);

int main(int argc, char** argv)
{
    if (argc > 3 && std::string(argv[1]) == "file") {
        Problem P;
        if (!load_file(argv[2], P)) { printf("FAIL cannot read %s\n", argv[2]); return 1; }
        return check(P, atoi(argv[3]));
    }
    const int dim = argc > 1 ? atoi(argv[1]) : 3;
    int nc[3] = {argc > 2 ? atoi(argv[2]) : 3, argc > 3 ? atoi(argv[3]) : 3, argc > 4 ? atoi(argv[4]) : 3};
    const int p = argc > 5 ? atoi(argv[5]) : 4;
    const int numbering = argc > 6 ? atoi(argv[6]) : 0;
    const int cap = argc > 7 ? atoi(argv[7]) : 1536;
    const unsigned seed = argc > 8 ? (unsigned)atoi(argv[8]) : 1u;
    const int wall = argc > 9 ? atoi(argv[9]) : 0;
    if (dim == 2) nc[2] = 0;
    int nd[3];
    for (int j = 0; j < 3; j++) nd[j] = j < dim ? nc[j] * p + 1 : 1;
    const int64_t n = (int64_t)nd[0] * nd[1] * nd[2];
    const int64_t stride = ((n + 31) / 32) * 32;
    std::mt19937_64 rng(seed);
    std::uniform_real_distribution<double> U(-1.0, 1.0);
    // numbering: user index of lexicographic point
    std::vector<int32_t> user_of_lex((size_t)n);
    std::iota(user_of_lex.begin(), user_of_lex.end(), 0);
    if (numbering) std::shuffle(user_of_lex.begin(), user_of_lex.end(), rng);
    nbgrid::Grid g;
    g.dim = dim; g.fe_order = p;
    for (int j = 0; j < 3; j++) g.n[j] = nd[j];
    g.nxp = (nd[0] + 1) & ~1;
    g.G = g.nxp * nd[1] * nd[2];
    g.gidx_of_int.assign((size_t)n, -1);
    for (int z = 0; z < nd[2]; z++) for (int y = 0; y < nd[1]; y++) for (int x = 0; x < nd[0]; x++)
        g.gidx_of_int[(size_t)user_of_lex[(size_t)(((int64_t)z * nd[1] + y) * nd[0] + x)]] = (int32_t)g.flat(x, y, z);
    // directions: all sign combinations with at least one non-zero component
    std::vector<std::array<int, 3>> dirs_e;
    for (int sz = (dim == 3 ? -1 : 0); sz <= (dim == 3 ? 1 : 0); sz++) for (int sy = -1; sy <= 1; sy++) for (int sx = -1; sx <= 1; sx++)
        if (sx || sy || sz) dirs_e.push_back(std::array<int, 3>{{sx, sy, sz}});
    const int ndir = (int)dirs_e.size();
    // 1D interpolation weights per (sign, position in cell, k): arbitrary numbers (the product structure is what matters)
    std::vector<double> w1((size_t)2 * (p + 1) * (p + 1));
    for (auto& v : w1) v = U(rng);
    auto track = [&](int axis, int X, int s, int* cols, double* wts) -> int {
        if (s == 0) { cols[0] = X; wts[0] = 1.0; return 1; }
        int c = X > 0 ? (X - 1) / p : 0, loc = X - c * p;
        if (s > 0 && loc == 0) c = (c + nc[axis] - 1) % nc[axis];       // departure towards -: leaves through the low face
        if (s < 0 && loc == p) c = (c + 1) % nc[axis];                  // departure towards +
        for (int k = 0; k <= p; k++) { cols[k] = c * p + k; wts[k] = w1[(size_t)(((s > 0) * (p + 1) + loc) * (p + 1) + k)]; }
        return p + 1;
    };
    std::vector<double> x((size_t)(ndir + 1) * stride);
    for (auto& v : x) v = U(rng);
    std::vector<nbdict::DirBuild> dirs((size_t)ndir);
    std::vector<std::vector<int64_t>> rowptr((size_t)ndir);
    std::vector<std::vector<int32_t>> col((size_t)ndir);       // user numbering, unsorted
    std::vector<std::vector<double>> val((size_t)ndir);
    std::vector<std::vector<int32_t>> wall_col((size_t)ndir);  // extra off-diagonal entries (population of direction (a+1) % ndir)
    std::vector<std::vector<double>> wall_val((size_t)ndir);
    std::vector<std::vector<int64_t>> wall_ptr((size_t)ndir);
    for (int a = 0; a < ndir; a++) {
        dirs[(size_t)a].init(n);
        std::vector<std::vector<int32_t>> rc((size_t)n);
        std::vector<std::vector<double>> rv((size_t)n);
        for (int Z = 0; Z < nd[2]; Z++) for (int Y = 0; Y < nd[1]; Y++) for (int X = 0; X < nd[0]; X++) {
            int cx[8], cy[8], cz[8];
            double wx[8], wy[8], wz[8];
            const int kx = track(0, X, dirs_e[(size_t)a][0], cx, wx), ky = track(1, Y, dirs_e[(size_t)a][1], cy, wy);
            const int kz = dim == 3 ? track(2, Z, dirs_e[(size_t)a][2], cz, wz) : (cz[0] = 0, wz[0] = 1.0, 1);
            const int32_t r = user_of_lex[(size_t)(((int64_t)Z * nd[1] + Y) * nd[0] + X)];
            for (int c = 0; c < kz; c++) for (int b = 0; b < ky; b++) for (int aa = 0; aa < kx; aa++) {
                rc[(size_t)r].push_back(user_of_lex[(size_t)(((int64_t)cz[c] * nd[1] + cy[b]) * nd[0] + cx[aa])]);
                rv[(size_t)r].push_back(wx[aa] * wy[b] * wz[c]);
            }
            if (wall && (r % 13) == 4 && rc[(size_t)r].size() > 2) { rc[(size_t)r].resize(rc[(size_t)r].size() - 2); rv[(size_t)r].resize(rv[(size_t)r].size() - 2); }
            // shuffled entry order: the library sorts by grid position
            std::vector<size_t> o(rc[(size_t)r].size());
            std::iota(o.begin(), o.end(), 0);
            std::shuffle(o.begin(), o.end(), rng);
            std::vector<int32_t> c2(o.size());
            std::vector<double> v2(o.size());
            for (size_t i = 0; i < o.size(); i++) { c2[i] = rc[(size_t)r][o[i]]; v2[i] = rv[(size_t)r][o[i]]; }
            rc[(size_t)r].swap(c2); rv[(size_t)r].swap(v2);
        }
        auto& rp = rowptr[(size_t)a];
        rp.push_back(0);
        wall_ptr[(size_t)a].push_back(0);
        for (int64_t r = 0; r < n; r++) {
            col[(size_t)a].insert(col[(size_t)a].end(), rc[(size_t)r].begin(), rc[(size_t)r].end());
            val[(size_t)a].insert(val[(size_t)a].end(), rv[(size_t)r].begin(), rv[(size_t)r].end());
            rp.push_back((int64_t)col[(size_t)a].size());
            if (wall && (r % 17) == 3) { wall_col[(size_t)a].push_back((int32_t)((r * 7 + 1) % n)); wall_val[(size_t)a].push_back(0.25); }
            wall_ptr[(size_t)a].push_back((int64_t)wall_col[(size_t)a].size());
        }
        // what nb200_upload_block_csr does: sort by grid position, then the dictionary
        std::vector<int32_t> sc;
        std::vector<double> sv;
        const int32_t* cp = col[(size_t)a].data();
        const double* vp = val[(size_t)a].data();
        if (nbgrid::sort_rows_by_grid(g.gidx_of_int, n, rp.data(), cp, vp, sc, sv)) { cp = sc.data(); vp = sv.data(); }
        const char* msg = "";
        if (!nbdict::add_block(dirs[(size_t)a], n, rp.data(), cp, vp, (int64_t)(a + 1) * stride, 0.0, 63, (1 << 26) - 1, &msg)) { printf("FAIL add_block %s\n", msg); return 1; }
        if (wall) {
            const int b = (a + 1) % ndir;
            if (!nbdict::add_block(dirs[(size_t)a], n, wall_ptr[(size_t)a].data(), wall_col[(size_t)a].data(), wall_val[(size_t)a].data(),
                                   (int64_t)(b + 1) * stride, 0.0, 63, (1 << 26) - 1, &msg)) { printf("FAIL add_block wall %s\n", msg); return 1; }
        }
    }
    Problem P;
    P.dim = dim; P.p = p; P.ndir = ndir; P.wall = wall; P.n = n; P.nloc = n; P.stride = stride; P.g = g;
    P.x.swap(x); P.dirs.swap(dirs); P.rowptr.swap(rowptr); P.col.swap(col); P.val.swap(val);
    P.wall_ptr.swap(wall_ptr); P.wall_col.swap(wall_col); P.wall_val.swap(wall_val);
    return check(P, cap);
}

static int check(Problem& P, int cap)
{
    const int64_t n = P.n, stride = P.stride;
    const int ndir = P.ndir, wall = P.wall;
    nbgrid::Grid& g = P.g;
    std::vector<double>& x = P.x;
    std::vector<nbdict::DirBuild>& dirs = P.dirs;
    auto &rowptr = P.rowptr, &wall_ptr = P.wall_ptr;
    auto &col = P.col, &wall_col = P.wall_col;
    auto &val = P.val, &wall_val = P.wall_val;
    for (auto& d : dirs) d.majority_class_first();      // as finalize_dict does
    nbgrid::Tables T;
    double pair_share = 1.0;
    const int max_k = 128;
    const double pair_tol = getenv("GRID_CHECK_TOL") ? 2.0 * atof(getenv("GRID_CHECK_TOL")) : 0.0;      // as finalize_dict: twice the value tolerance
    if (!nbgrid::build(dirs, g, n, stride, 128, cap, max_k, 63, T, pair_tol)) { printf("INFEASIBLE\n"); return 0; }
    // grid copy of x
    const int64_t gstride = (g.G + 31) / 32 * 32;
    std::vector<double> xg((size_t)(ndir + 1) * gstride, 0.0);
    for (int q = 0; q <= ndir; q++) for (int64_t i = 0; i < P.nloc; i++) xg[(size_t)(q * gstride + g.gidx_of_int[(size_t)i])] = x[(size_t)(q * stride + i)];
    double max_err = 0.0, max_ref = 0.0;
    int64_t checked = 0, seen_rows = 0;
    std::vector<double> xs((size_t)cap);
    std::vector<uint8_t> row_seen((size_t)n, 0);
    for (int64_t b = 0; b < T.n_tiles; b++) {
        int next_dir = 0;
        for (int t = 0; t < 128; t++) {
            const int32_t r = T.tile_row[(size_t)(b * 128 + t)];
            if (r >= 0) {
                if (row_seen[(size_t)r]) { printf("FAIL row %d in two tiles\n", r); return 1; }
                row_seen[(size_t)r] = 1; seen_rows++;
                if (T.tile_gidx[(size_t)(b * 128 + t)] != g.gidx_of_int[(size_t)r]) { printf("FAIL tile_gidx\n"); return 1; }
            }
        }
        for (int pi = T.tile_pass[(size_t)b]; pi < T.tile_pass[(size_t)b + 1]; pi++) {
            const auto& ps = T.passes[(size_t)pi];
            if (ps.a0 != next_dir) { printf("FAIL pass table\n"); return 1; }
            next_dir = ps.a1;
            std::fill(xs.begin(), xs.end(), 1e300);      // poison: values outside the boxes must never be used
            int64_t bytes = 0;
            for (int bi = 0; bi < ps.n_box; bi++) {
                const auto& bx = T.boxes[(size_t)(ps.box_begin + bi)];
                const int32_t* bd = T.box_dims.data() + (size_t)bx.dir * 3;
                if (bx.smem_off % 16 || bx.smem_off + (int64_t)bd[0] * bd[1] * bd[2] > cap || bx.dir < ps.a0 || bx.dir >= ps.a1) { printf("FAIL box placement\n"); return 1; }
                for (int z = 0; z < bd[2]; z++) for (int y = 0; y < bd[1]; y++) for (int xx = 0; xx < bd[0]; xx++) {
                    const int gx = bx.x + xx, gy = bx.y + y, gz = bx.z + z;
                    double v = 0.0;
                    if (gx >= 0 && gx < g.nxp && gy >= 0 && gy < g.n[1] && gz >= 0 && gz < g.n[2]) v = xg[(size_t)((bx.dir + 1) * gstride + g.flat(gx, gy, gz))];
                    xs[(size_t)(bx.smem_off + (z * bd[1] + y) * bd[0] + xx)] = v;
                }
                bytes += (int64_t)bd[0] * bd[1] * bd[2] * 8;
                if (bx.x & 1) { printf("FAIL odd box origin in x (TMA needs 16-byte aligned inner coordinates)\n"); return 1; }
            }
            if (bytes != ps.bytes) { printf("FAIL pass bytes %lld != %d\n", (long long)bytes, ps.bytes); return 1; }
            for (int a = ps.a0; a < ps.a1; a++)
                for (int t = 0; t < 128; t++) {
                    const int32_t r = T.tile_row[(size_t)(b * 128 + t)];
                    const uint32_t dx = (uint32_t)T.desc_x[(size_t)a * T.desc_stride + (size_t)(b * 128 + t)];
                    if (r < 0) { if ((dx >> 16) != 63) { printf("FAIL idle thread descriptor\n"); return 1; } continue; }
                    const auto& d = dirs[(size_t)a];
                    double ref = 0.0, aref = 0.0;
                    for (int64_t k = rowptr[(size_t)a][(size_t)r]; k < rowptr[(size_t)a][(size_t)r + 1]; k++) {
                        const double tt = val[(size_t)a][(size_t)k] * x[(size_t)((int64_t)(a + 1) * stride + col[(size_t)a][(size_t)k])];
                        ref += tt; aref += std::fabs(tt);
                    }
                    if (!P.extra.empty()) for (const ExtraBlock& B : P.extra[(size_t)a])
                        for (int64_t k = B.rowptr[(size_t)r]; k < B.rowptr[(size_t)r + 1]; k++) {
                            const double tt = B.val[(size_t)k] * x[(size_t)((int64_t)(B.bj + 1) * stride + B.col[(size_t)k])];
                            ref += tt; aref += std::fabs(tt);
                        }
                    if (wall) for (int64_t k = wall_ptr[(size_t)a][(size_t)r]; k < wall_ptr[(size_t)a][(size_t)r + 1]; k++) {
                        const double tt = wall_val[(size_t)a][(size_t)k] * x[(size_t)((int64_t)((a + 1) % ndir + 1) * stride + wall_col[(size_t)a][(size_t)k])];
                        ref += tt; aref += std::fabs(tt);
                    }
                    double got = 0.0;
                    const int ci = d.row_cls[(size_t)r];
                    if (dx >> 31) {          // generic: dictionary list
                        const auto& C = d.cls[(size_t)ci];
                        const double* W = C.pats.data() + (size_t)d.row_pat[(size_t)r] * C.K;
                        const int32_t* L = C.lists.data() + (size_t)d.row_lst[(size_t)r] * C.K;
                        for (int k = 0; k < C.K; k++) got += W[k] * x[(size_t)L[k]];
                    } else if ((dx >> 16) == 0) {
                        if (ci != 0) { printf("FAIL box row of class %d\n", ci); return 1; }
                        const auto& C = d.cls[0];
                        const double* W = C.pats.data() + (size_t)T.desc_y[(size_t)a * T.desc_stride + (size_t)(b * 128 + t)] * C.K;
                        for (int k = 0; k < C.K; k++) got += W[k] * xs[(size_t)((dx & 0xffffu) + T.off_table[(size_t)a * max_k + k])];
                    } else if (ci >= 0) { printf("FAIL non-empty row with the empty descriptor\n"); return 1; }
                    max_err = std::max(max_err, std::fabs(got - ref));
                    max_ref = std::max(max_ref, aref);
                    checked++;
                }
        }
        if (next_dir != ndir) { printf("FAIL passes do not cover all directions\n"); return 1; }
    }
    {   // how many row pairs (slots t and t + 64 of a tile) take the paired product: both class-0 box rows with the same pattern id
        int64_t pairs = 0, paired = 0;
        for (int a = 0; a < ndir; a++)
            for (int64_t b = 0; b < T.n_tiles; b++)
                for (int t = 0; t < 64; t++) {
                    const size_t s0 = (size_t)(b * 128 + t), s1 = s0 + 64;
                    if (T.tile_row[s0] < 0 || T.tile_row[s1] < 0) continue;
                    pairs++;
                    const uint32_t x0 = (uint32_t)T.desc_x[(size_t)a * T.desc_stride + s0], x1 = (uint32_t)T.desc_x[(size_t)a * T.desc_stride + s1];
                    if (((x0 | x1) >> 16) == 0 && T.desc_y[(size_t)a * T.desc_stride + s0] == T.desc_y[(size_t)a * T.desc_stride + s1]) paired++;
                }
        pair_share = pairs ? (double)paired / (double)pairs : 1.0;
    }
    if (!T.pair_hist.empty()) {
        std::vector<double> h = T.pair_hist;
        std::sort(h.begin(), h.end());
        printf("  mismatched pairs %zu: entry difference min %.2e median %.2e p90 %.2e p99 %.2e max %.2e\n", h.size(), h.front(), h[h.size() / 2],
               h[h.size() * 9 / 10], h[h.size() * 99 / 100], h.back());
    }
    if (getenv("GRID_CHECK_VERBOSE")) {       // per direction: rows from boxes / from their lists, class-0 row length, box shape
        for (int a = 0; a < ndir; a++) {
            int64_t gen = 0, box = 0;
            for (int64_t sl = 0; sl < T.n_tiles * 128; sl++) {
                if (T.tile_row[(size_t)sl] < 0) continue;
                const uint32_t dx = (uint32_t)T.desc_x[(size_t)a * T.desc_stride + (size_t)sl];
                if (dx >> 31) gen++; else if ((dx >> 16) == 0) box++;
            }
            printf("  dir %2d: K0=%d classes=%zu box=%lld generic=%lld box_dims=%dx%dx%d\n", a, dirs[(size_t)a].cls.empty() ? 0 : dirs[(size_t)a].cls[0].K,
                   dirs[(size_t)a].cls.size(), (long long)box, (long long)gen, T.box_dims[(size_t)a * 3], T.box_dims[(size_t)a * 3 + 1], T.box_dims[(size_t)a * 3 + 2]);
        }
    }
    if (seen_rows != n) { printf("FAIL %lld of %lld rows in tiles\n", (long long)seen_rows, (long long)n); return 1; }
    // box stores: a flagged half-tile written as one box (out-of-range points clipped) must put every row of the half at its own
    // grid point and touch no grid point that holds another DoF
    int64_t store_halves = 0, store_rows = 0;
    {
        std::vector<int32_t> dof_at((size_t)g.G, -1);
        for (size_t i = 0; i < g.gidx_of_int.size(); i++) dof_at[(size_t)g.gidx_of_int[i]] = (int32_t)i;
        const int* hd = T.half_dims;
        if (T.tile_store.size() != (size_t)T.n_tiles * 4) { printf("FAIL tile_store size\n"); return 1; }
        for (int64_t b = 0; b < T.n_tiles; b++)
            for (int hf = 0; hf < 2; hf++) {
                const int16_t* ts = T.tile_store.data() + (size_t)b * 4;
                if (!((ts[3] >> hf) & 1)) continue;
                store_halves++;
                const int bx = ts[0] + hf * hd[0];
                if ((bx & 1) || (hd[0] * 8) % 16) { printf("FAIL box store alignment\n"); return 1; }
                for (int tt = 0; tt < hd[0] * hd[1] * hd[2]; tt++) {
                    const int c[3] = {bx + tt % hd[0], ts[1] + (tt / hd[0]) % hd[1], ts[2] + tt / (hd[0] * hd[1])};
                    const int32_t r = T.tile_row[(size_t)(b * 128 + hf * 64 + tt)];
                    const bool inside = c[0] >= 0 && c[0] < g.nxp && c[1] >= 0 && c[1] < g.n[1] && c[2] >= 0 && c[2] < g.n[2];
                    if (!inside) { if (r >= 0) { printf("FAIL row outside the grid in a box store\n"); return 1; } continue; }
                    const int64_t f = g.flat(c[0], c[1], c[2]);
                    if (r >= 0) { if (g.gidx_of_int[(size_t)r] != f) { printf("FAIL box store puts a row elsewhere\n"); return 1; } store_rows++; }
                    else if (dof_at[(size_t)f] >= 0) { printf("FAIL box store overwrites another DoF\n"); return 1; }
                }
                for (int tt = hd[0] * hd[1] * hd[2]; tt < 64; tt++)
                    if (T.tile_row[(size_t)(b * 128 + hf * 64 + tt)] >= 0) { printf("FAIL row beyond the half-tile box\n"); return 1; }
            }
    }
    if (!(max_err <= 1e-13 * max_ref)) { printf("FAIL max_err %g (scale %g)\n", max_err, max_ref); return 1; }
    printf("OK rows=%lld checked=%lld tiles=%lld boxes=%lld passes=%zu box_rows=%lld generic=%lld max_pass=%lld err=%.2e store_halves=%lld store_rows=%lld paired=%.4f unified=%lld\n", (long long)n, (long long)checked,
           (long long)T.n_tiles, (long long)T.total_boxes, T.passes.size(), (long long)T.grid_rows, (long long)T.generic_rows, (long long)T.max_pass_doubles, max_err,
           (long long)store_halves, (long long)store_rows, pair_share, (long long)T.pairs_unified);
    return 0;
}
