// Host-only check of the NB_FMT_STAGED driving tables (natrium_b200/csrc/dict_build.h): builds the dictionary and
// the staging tables for a synthetic block-structured matrix, replays the kernel's access pattern on the CPU
// (stage_col -> xs, row = sum_k W[pattern][k] * xs[off + k]) and compares with the plain CSR product.
// Usage: staging_check <n_rows> <n_dirs> <rows_per_cell> <K> <cap> <seed>; prints "OK ..." or "FAIL ...".
#include <cstdio>
#include <cstdlib>
#include <random>
#include "../../natrium_b200/csrc/dict_build.h"

int main(int argc, char** argv)
{
    const int64_t n = argc > 1 ? atoll(argv[1]) : 1000;
    const int nd = argc > 2 ? atoi(argv[2]) : 4;
    const int rpc = argc > 3 ? atoi(argv[3]) : 16;
    const int K = argc > 4 ? atoi(argv[4]) : 9;
    const int cap = argc > 5 ? atoi(argv[5]) : 4096;
    const unsigned seed = argc > 6 ? (unsigned)atoi(argv[6]) : 1u;
    const int64_t stride = ((n + 31) / 32) * 32;
    std::mt19937_64 rng(seed);
    std::uniform_real_distribution<double> U(-1.0, 1.0);
    std::vector<double> x((size_t)(nd + 1) * stride);
    for (auto& v : x) v = U(rng);
    std::vector<nbdict::DirBuild> dirs((size_t)nd);
    std::vector<std::vector<int64_t>> rowptr((size_t)nd);
    std::vector<std::vector<int32_t>> col((size_t)nd);
    std::vector<std::vector<double>> val((size_t)nd);
    const int64_t n_cells = (n + rpc - 1) / rpc;
    for (int a = 0; a < nd; a++) {
        dirs[(size_t)a].init(n);
        // direction a: row r reads the K columns of "cell" (r / rpc + shift) with the weights of position r % rpc;
        // every 7th row of odd directions is shorter (second row-length class), every 11th row is empty
        std::vector<std::vector<int32_t>> cell_cols((size_t)n_cells);
        for (auto& cc : cell_cols) { cc.resize((size_t)K); for (auto& v : cc) v = (int32_t)(rng() % (uint64_t)n); }
        std::vector<std::vector<double>> pos_w((size_t)rpc);
        for (auto& w : pos_w) { w.resize((size_t)K); for (auto& v : w) v = U(rng); }
        auto& rp = rowptr[(size_t)a]; auto& cc = col[(size_t)a]; auto& vv = val[(size_t)a];
        rp.push_back(0);
        for (int64_t r = 0; r < n; r++) {
            const int64_t cell = (r / rpc + a) % n_cells;
            int k_row = K;
            if ((a & 1) && r % 7 == 3) k_row = K > 2 ? K - 2 : K;
            if (r % 11 == 5) k_row = 0;
            for (int k = 0; k < k_row; k++) { cc.push_back(cell_cols[(size_t)cell][(size_t)k]); vv.push_back(pos_w[(size_t)(r % rpc)][(size_t)k]); }
            rp.push_back((int64_t)cc.size());
        }
        const char* msg = "";
        if (!nbdict::add_block(dirs[(size_t)a], n, rp.data(), cc.data(), vv.data(), (int64_t)(a + 1) * stride, 0.0, 63, (1 << 26) - 1, &msg)) {
            printf("FAIL add_block: %s\n", msg);
            return 1;
        }
    }
    const int64_t desc_stride = stride;
    nbdict::StagingBuild SB;
    const bool ok = nbdict::build_staging(dirs, n, desc_stride, 128, cap, 63, SB);
    if (!ok) { printf("INFEASIBLE\n"); return 0; }
    const int64_t n_cta = (n + 127) / 128;
    double max_err = 0.0;
    std::vector<double> xs((size_t)cap);
    int64_t checked = 0;
    for (int64_t b = 0; b < n_cta; b++) {
        int next_dir = 0;
        for (int p = SB.cta_ptr[(size_t)b]; p < SB.cta_ptr[(size_t)b + 1]; p++) {
            const auto& ps = SB.passes[(size_t)p];
            if (ps.a0 != next_dir || ps.count > cap || ps.count < 0) { printf("FAIL pass table (cta %lld)\n", (long long)b); return 1; }
            next_dir = ps.a1;
            for (int e = 0; e < ps.count; e++) xs[(size_t)e] = x[(size_t)SB.stage_col[(size_t)(ps.begin + e)]];
            for (int a = ps.a0; a < ps.a1; a++)
                for (int64_t r = b * 128; r < std::min<int64_t>(n, b * 128 + 128); r++) {
                    const uint32_t dx = (uint32_t)SB.sdesc_x[(size_t)a * desc_stride + r];
                    const int ci = (int)(dx >> 16);
                    const auto& d = dirs[(size_t)a];
                    double ref = 0.0;
                    for (int64_t k = rowptr[(size_t)a][(size_t)r]; k < rowptr[(size_t)a][(size_t)r + 1]; k++)
                        ref += val[(size_t)a][(size_t)k] * x[(size_t)((int64_t)(a + 1) * stride + col[(size_t)a][(size_t)k])];
                    double got = 0.0;
                    if (d.row_cls[(size_t)r] < 0) {
                        if (ci != 63) { printf("FAIL empty row class\n"); return 1; }
                    } else {
                        if (ci != d.row_cls[(size_t)r]) { printf("FAIL class id\n"); return 1; }
                        const auto& C = d.cls[(size_t)ci];
                        const double* W = C.pats.data() + (size_t)d.row_pat[(size_t)r] * C.K;
                        const int off = (int)(dx & 0xffffu);
                        if (off + C.K > ps.count) { printf("FAIL offset out of pass\n"); return 1; }
                        for (int k = 0; k < C.K; k++) got += W[k] * xs[(size_t)(off + k)];
                    }
                    max_err = std::max(max_err, std::fabs(got - ref));
                    checked++;
                }
        }
        if (next_dir != nd) { printf("FAIL passes do not cover all directions\n"); return 1; }
    }
    if (max_err != 0.0) { printf("FAIL max_err %g\n", max_err); return 1; }
    printf("OK rows=%lld checked=%lld passes=%zu staged=%zu max_pass=%lld\n", (long long)n, (long long)checked, SB.passes.size(), SB.stage_col.size(), (long long)SB.max_pass_count);
    return 0;
}
