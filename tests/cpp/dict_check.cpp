// Host-only checks of the dictionary builder (natrium_b200/csrc/dict_build.h): value tolerance semantics and the
// bounded number of row-length classes.  Prints "OK ..." or "FAIL ...".
#include <cstdio>
#include <cstdlib>
#include <random>
#include "../../natrium_b200/csrc/dict_build.h"

static double replay_err(const nbdict::DirBuild& d, int64_t n, const std::vector<int64_t>& rp, const std::vector<int32_t>& col,
                         const std::vector<double>& val, const std::vector<double>& x, int64_t col_base)
{
    double worst = 0.0;
    for (int64_t r = 0; r < n; r++) {
        double ref = 0.0, got = 0.0;
        for (int64_t k = rp[(size_t)r]; k < rp[(size_t)r + 1]; k++) ref += val[(size_t)k] * x[(size_t)(col_base + col[(size_t)k])];
        const int ci = d.row_cls[(size_t)r];
        if (ci >= 0) {
            const auto& C = d.cls[(size_t)ci];
            const int32_t* L = C.lists.data() + (size_t)d.row_lst[(size_t)r] * C.K;
            const double* W = C.pats.data() + (size_t)d.row_pat[(size_t)r] * C.K;
            for (int k = 0; k < C.K; k++) got += W[k] * x[(size_t)L[k]];
        }
        worst = std::max(worst, std::fabs(got - ref));
    }
    return worst;
}

int main()
{
    std::mt19937_64 rng(7);
    std::uniform_real_distribution<double> U(0.1, 1.0);
    // ---- 1. tolerance: 64 "positions" x 500 "cells"; the values of a position differ between cells by round-off noise
    {
        const int64_t n = 32000, stride = n;
        const int K = 9, npos = 64;
        std::vector<std::vector<double>> base((size_t)npos, std::vector<double>((size_t)K));
        for (auto& b : base) for (auto& v : b) v = U(rng);
        std::vector<int64_t> rp{0};
        std::vector<int32_t> col;
        std::vector<double> val;
        std::uniform_real_distribution<double> noise(-4e-16, 4e-16);
        for (int64_t r = 0; r < n; r++) {
            for (int k = 0; k < K; k++) { col.push_back((int32_t)((r / npos * 7 + k * 13) % n)); val.push_back(base[(size_t)(r % npos)][(size_t)k] + noise(rng)); }
            rp.push_back((int64_t)col.size());
        }
        std::vector<double> x((size_t)(2 * stride));
        for (auto& v : x) v = U(rng);
        const char* msg = "";
        for (double tol : {0.0, 1e-14}) {
            nbdict::DirBuild d;
            d.init(n);
            if (!nbdict::add_block(d, n, rp.data(), col.data(), val.data(), stride, tol, 63, (1 << 26) - 1, &msg)) { printf("FAIL add_block %s\n", msg); return 1; }
            const int64_t np = d.cls[0].n_pats(), nl = d.cls[0].n_lists();
            const double err = replay_err(d, n, rp, col, val, x, stride);
            if (tol == 0.0 && (np < n / 2 || err != 0.0)) { printf("FAIL tol=0: patterns %lld err %g (bitwise dedup must keep noisy rows apart and be exact)\n", (long long)np, err); return 1; }
            if (tol > 0.0 && (np != npos || err > K * tol * 1.0)) { printf("FAIL tol=1e-14: patterns %lld err %g\n", (long long)np, err); return 1; }
            if (nl != (n + npos - 1) / npos) { printf("FAIL lists %lld\n", (long long)nl); return 1; }
        }
        // values 1e-9 apart are different rows even with the default tolerance
        nbdict::DirBuild d;
        d.init(2);
        std::vector<int64_t> rp2{0, 2, 4};
        std::vector<int32_t> c2{0, 1, 0, 1};
        std::vector<double> v2{0.5, 0.5, 0.5 + 1e-9, 0.5 - 1e-9};
        nbdict::add_block(d, 2, rp2.data(), c2.data(), v2.data(), 0, 1e-14, 63, (1 << 26) - 1, &msg);
        if (d.cls[0].n_pats() != 2 || d.cls[0].n_lists() != 1) { printf("FAIL 1e-9 apart merged\n"); return 1; }
    }
    // ---- 2. more distinct row lengths than exact classes: power-of-two classes with zero-weight padding take over
    {
        const int64_t n = 300, stride = 512;
        std::vector<int64_t> rp{0};
        std::vector<int32_t> col;
        std::vector<double> val;
        for (int64_t r = 0; r < n; r++) {
            const int K = 1 + (int)(r % 120);            // 120 distinct lengths
            for (int k = 0; k < K; k++) { col.push_back((int32_t)((r * 31 + k * 7) % n)); val.push_back(U(rng)); }
            rp.push_back((int64_t)col.size());
        }
        std::vector<double> x((size_t)(2 * stride));
        for (auto& v : x) v = U(rng);
        nbdict::DirBuild d;
        d.init(n);
        const char* msg = "";
        if (!nbdict::add_block(d, n, rp.data(), col.data(), val.data(), stride, 0.0, 63, (1 << 26) - 1, &msg)) { printf("FAIL add_block %s\n", msg); return 1; }
        if ((int)d.cls.size() > 63) { printf("FAIL %zu classes\n", d.cls.size()); return 1; }
        const double err = replay_err(d, n, rp, col, val, x, stride);
        if (err > 1e-13) { printf("FAIL padded classes err %g\n", err); return 1; }
        // a second block of the same block-row is appended to the rows (wall bounce: two blocks per row)
        std::vector<int64_t> rpb{0};
        std::vector<int32_t> colb;
        std::vector<double> valb;
        for (int64_t r = 0; r < n; r++) {
            if (r % 3 == 0) { colb.push_back((int32_t)(r % n)); valb.push_back(0.25); }
            rpb.push_back((int64_t)colb.size());
        }
        if (!nbdict::add_block(d, n, rpb.data(), colb.data(), valb.data(), 0, 0.0, 63, (1 << 26) - 1, &msg)) { printf("FAIL add_block 2 %s\n", msg); return 1; }
        double worst = 0.0;
        for (int64_t r = 0; r < n; r++) {
            double ref = 0.0, got = 0.0;
            for (int64_t k = rp[(size_t)r]; k < rp[(size_t)r + 1]; k++) ref += val[(size_t)k] * x[(size_t)(stride + col[(size_t)k])];
            for (int64_t k = rpb[(size_t)r]; k < rpb[(size_t)r + 1]; k++) ref += valb[(size_t)k] * x[(size_t)colb[(size_t)k]];
            const auto& C = d.cls[(size_t)d.row_cls[(size_t)r]];
            const int32_t* L = C.lists.data() + (size_t)d.row_lst[(size_t)r] * C.K;
            const double* W = C.pats.data() + (size_t)d.row_pat[(size_t)r] * C.K;
            for (int k = 0; k < C.K; k++) got += W[k] * x[(size_t)L[k]];
            worst = std::max(worst, std::fabs(got - ref));
        }
        if (worst > 1e-13) { printf("FAIL two-block rows err %g\n", worst); return 1; }
    }
    printf("OK\n");
    return 0;
}
