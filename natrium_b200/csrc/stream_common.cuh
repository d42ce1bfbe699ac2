// stream_common.cuh -- the device formats of the streaming matrix (ELL, dictionary, staged dictionary) and their row products.
//
//   NB_FMT_ELL   warp-sliced ELL: 8 B value + 4 B index per stored entry, streamed from HBM.
//   NB_FMT_DICT  dictionary format: every row is a (column-list id, weight-pattern id) pair; lists
//                (list-major) and patterns (k-major) live in pools, one pool per row length K ("class").  Rows whose
//                departure point lies in the same source cell share a column list; rows with the
//                same position relative to their cell share a weight pattern (up to the dedup
//                tolerance), so on a regular mesh the matrix shrinks from 12 B/nnz to 8 B/row and the
//                kernel is bound by the populations, not by the matrix.  Without any sharing the
//                pools degenerate to a column-major ELL, so the format is always valid.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

//   NB_FMT_STAGED  the dictionary format driven CTA by CTA: for every block of NB_CTA_ROWS rows the builder
//                lists, pass by pass, the distinct column lists those rows read (stage_col: one flat population
//                index per staged support value).  The CTA copies them into shared memory with independent,
//                coalesced loads, then every row takes its K support values from there (xs[off + k]) and its
//                weights from the pattern pool.  A pass covers a range of directions whose staged values fit
//                NB_STAGE_CAP doubles.  Rows that read the same source cell share the staged copy, so the
//                gather traffic per CTA drops from rows x K to lists x K loads and the dependent
//                descriptor -> list -> value chain of NB_FMT_DICT disappears from the inner loop.
//   NB_FMT_GRID    the dictionary format driven tile by tile over a lexicographic grid copy of the populations
//                (nb200_set_dof_grid): the support values of a tile of rows are whole boxes of that grid and arrive by TMA
//                tensor copies (cp.async.bulk.tensor -> mbarrier), double buffered; see grid_build.h.
enum { NB_FMT_ELL = 0, NB_FMT_DICT = 1, NB_FMT_STAGED = 2, NB_FMT_GRID = 3 };

#define NB_CTA_ROWS 128
// Pass capacity in doubles per distribution.  Deliberately small: with 5 CTAs per SM the staged values take
// 5 x 16 KB (+ 5 x 19.5 KB of result tiles for Q = 19), which leaves the L1 large enough to keep the weight patterns
// resident (measured: 1536 / 2048 / 3072 doubles -> 0.743 / 0.729 / 0.875 ms per step on configuration 2).
#ifndef NB_STAGE_CAP
#define NB_STAGE_CAP 2048          // f only
#endif
#ifndef NB_STAGE_CAP_FG
#define NB_STAGE_CAP_FG 2048       // f and g staged together (two arrays of this size)
#endif
#define NB_MAX_DIRS 44             // Q - 1 of the largest stencil on the path (D3Q45)
// grid kernels: capacity of ONE staging buffer in doubles per distribution (two buffers: the TMA copies of the next
// pass land while the current one is multiplied), and the longest row the offset table holds ((p+1)^3 for p <= 4)
#ifndef NB_GRID_CAP
#define NB_GRID_CAP 1536
#endif
#ifndef NB_GRID_CAP_FG
#define NB_GRID_CAP_FG 1280        // f + g, large stencils (D3Q45: a 3-d box of 14 x 9 x 9 values per distribution)
#endif
#ifndef NB_GRID_CAP_FGF
#define NB_GRID_CAP_FGF 512        // f + g, stencils that run the fused kernel (Q <= 25, D2Q25H: small 2-d boxes)
#endif
// capacity of one staging buffer per distribution for a stencil with Q directions and n_rhs distributions
#ifndef NB_GRID_CAP_MULTI
#define NB_GRID_CAP_MULTI NB_GRID_CAP   // f only, several ranks with the overlapped exchange.  1024 (45 KB of an SM's shared memory
                                         // left for the exchange kernels) was measured on 2 GPUs: 0.639 against 0.602 ms/step with 1536
                                         // -- the extra passes cost more than the room gives
#endif
#define NB_GRID_CAP_OF(Q, n_rhs) ((n_rhs) == 2 ? ((Q) <= 25 ? NB_GRID_CAP_FGF : NB_GRID_CAP_FG) : NB_GRID_CAP)
#define NB_GRID_MAXK 128
// row lengths with a compile-time instance of the pair product: 3-d units add the lengths of three moving axes
#ifndef NB_GRID_K3D
#if defined(NB_D) && NB_D == 3
#define NB_GRID_K3D 1
#else
#define NB_GRID_K3D 0
#endif
#endif

#define NB_CLS_BITS 6
#define NB_MAX_CLS (1 << NB_CLS_BITS)           // row-length classes per direction
#define NB_PAT_BITS (32 - NB_CLS_BITS)
#define NB_PAT_MASK ((1u << NB_PAT_BITS) - 1u)

// Weight pool layout of a class: pairs of consecutive entries of a pattern sit next to each other,
// W[((k >> 1) * P + pattern) * 2 + (k & 1)], k-pair-major: the lanes of a warp hold different (neighbouring)
// patterns, so one 16-byte load per lane fetches two weights from one or two adjacent lines.  Odd K is padded
// with a zero weight.
__host__ __device__ __forceinline__ int64_t nb_w_off(int k, int64_t P) { return (int64_t)(k >> 1) * (2 * P) + (k & 1); }

// one row-length class of one direction
struct NbDirClass {
    const double* __restrict__ W;     // [(K+1)/2][P][2]  weight patterns, see nb_w_off
    const int32_t* __restrict__ L;    // [n_lists][NL] column lists (flat population index), list-major, 16-byte aligned rows
    int32_t K;
    int32_t streamed;                 // pools too large to stay cached: read with evict-first
    int64_t P, NL;                    // P: padded pattern count (k-major pitch); NL: list pitch = K rounded up to 4
};

// one staging pass of one CTA: directions [a0, a1), `count` staged values starting at stage_col[begin]
struct NbStagePass {
    int64_t begin;
    int32_t count;
    int16_t a0, a1;
};

// one staging pass of one tile (grid kernels): directions [a0, a1), boxes [box_begin, box_begin + n_box)
struct NbGridPass {
    int32_t box_begin;
    int16_t n_box, a0, a1, pad;
    int32_t bytes;          // TMA bytes of the pass for one distribution
};
// one TMA copy: box of direction dir's population with origin (x, y, z) -> staging buffer + smem_off
struct NbGridBox {
    int16_t x, y, z, dir;
    int32_t smem_off;
};

struct StreamArgs {
    // NB_FMT_ELL
    const double* __restrict__ ell_val;
    const int32_t* __restrict__ ell_idx;
    const int64_t* __restrict__ slice_off;   // [(Q-1)][n_slices+1]
    // NB_FMT_DICT
    const int2* __restrict__ desc;           // [(Q-1)][desc_stride]: x = list id, y = class << 26 | pattern id
    const NbDirClass* __restrict__ cls;      // [(Q-1)][NB_MAX_CLS]
    int64_t desc_stride;
    // NB_FMT_STAGED (shares cls / desc_stride with NB_FMT_DICT)
    const int2* __restrict__ sdesc;          // [(Q-1)][desc_stride]: x = offset into the pass's staged values | class << 16, y = pattern id
    const int32_t* __restrict__ stage_col;   // flat population index of every staged support value, pass after pass
    const struct NbStagePass* __restrict__ stage_pass;
    const int32_t* __restrict__ stage_cta;   // [n_cta + 1] first pass of every CTA
    const int32_t* __restrict__ cta_map;     // launch over a subset of the CTAs (interior / boundary lists); null = all
    // class 0 of every direction (the row length almost every row has) is described right here in the kernel
    // arguments, so the common case needs no dependent table load between descriptor and weights
    const double* c0_W[NB_MAX_DIRS];
    int32_t c0_P[NB_MAX_DIRS];
    int32_t c0_K[NB_MAX_DIRS];               // K | streamed << 30
    int64_t n_slices;
    int64_t n_owned;
    int64_t stride;
    // NB_FMT_GRID (shares cls / desc / c0_* with the dictionary; sdesc holds the grid descriptors, stage_cta the first pass
    // of every tile)
    const int32_t* __restrict__ tile_row;    // [n_tiles][NB_CTA_ROWS] canonical row of every tile thread, -1 = idle
    const int32_t* __restrict__ tile_gidx;   // [n_tiles][NB_CTA_ROWS] flat index in the grid copy
    const NbGridPass* __restrict__ gpass;
    const NbGridBox* __restrict__ gbox;
    const void* tmap_f;                      // [(Q-1)] CUtensorMap of population a+1 in the current grid copy of f
    const void* tmap_g;                      //         ... of g
    int64_t gstride;                         // population pitch of the grid copies
    int64_t gdesc_stride;                    // pitch of the grid descriptors (sdesc) = n_tiles * NB_CTA_ROWS
    const short4* __restrict__ tile_store;   // [n_tiles] grid origin of the tile's first half + flags: bit h = half h is written to
                                             // the grid copy as one box (TMA store); null = per-thread stores only
    const void* tmap_out_f;                  // [Q] CUtensorMap (box = one half-tile) of every population in the NEXT grid copy of f
    int half_x;                              // x extent of a half-tile
    int grid_cap;                            // values per staging buffer of the fused f kernel (<= NB_GRID_CAP; the tables were
                                             // built for it): smaller when other kernels must fit next to it on an SM
};

// one ELL row dot product for 1 or 2 right-hand sides (f and g share the matrix pass)
template <int NRHS>
__device__ __forceinline__ void nb_row_dot_ell(const StreamArgs& A, int alpha_m1, int64_t slice, int lane,
                                               const double* __restrict__ x0, const double* __restrict__ x1,
                                               double& y0, double& y1)
{
    const int64_t* so = A.slice_off + (int64_t)alpha_m1 * (A.n_slices + 1) + slice;
    const int64_t off = __ldg(so);
    const int w = (int)((__ldg(so + 1) - off) >> 5);
    const double* __restrict__ v = A.ell_val + off + lane;
    const int32_t* __restrict__ ix = A.ell_idx + off + lane;
    double a0 = 0.0, a1 = 0.0;
    int k = 0;
    // rows on the path have 5, 25 or 125 entries ((p+1)^k): batches of 5 keep 5 value/index
    // loads and 5 gathers in flight per thread without a remainder loop in the common case
    for (; k + 5 <= w; k += 5) {
        double vv[5];
        int32_t ii[5];
#pragma unroll
        for (int j = 0; j < 5; j++) {
            vv[j] = __ldcs(v + (int64_t)(k + j) * 32);     // streamed once: evict-first
            ii[j] = __ldcs(ix + (int64_t)(k + j) * 32);
        }
#pragma unroll
        for (int j = 0; j < 5; j++) {
            a0 += vv[j] * __ldg(x0 + ii[j]);
            if (NRHS == 2) a1 += vv[j] * __ldg(x1 + ii[j]);
        }
    }
    for (; k < w; k++) {
        const double vv = __ldcs(v + (int64_t)k * 32);
        const int32_t ii = __ldcs(ix + (int64_t)k * 32);
        a0 += vv * __ldg(x0 + ii);
        if (NRHS == 2) a1 += vv * __ldg(x1 + ii);
    }
    y0 = a0;
    y1 = a1;
}

// Dictionary row product, one row per lane.  The column list of a row is contiguous (list-major pool,
// padded to a multiple of 4 entries) and is read with 16-byte loads; lanes whose rows read the same
// source cell read the same addresses (one L1 wavefront).  Weight patterns are k-major: lanes hold
// different (neighbouring) pattern ids, so one k is one or two adjacent lines.  With a cell-blocked DoF
// order all three streams (list, weights, gathered support values) are L1/L2 hits.
// Summation order per row is k = 0..K-1, exactly as stored (and as the CSR reference loop).
template <int NRHS, bool STREAMED>
__device__ __forceinline__ void nb_dict_accumulate(const double* __restrict__ W, const int32_t* __restrict__ L,
                                                   int K, int64_t P, const double* __restrict__ x0,
                                                   const double* __restrict__ x1, double& a0, double& a1)
{
    const int4* __restrict__ L4 = reinterpret_cast<const int4*>(L);
    int k = 0;
    for (; k + 8 <= K; k += 8) {
        const int4 i0 = STREAMED ? __ldcs(L4 + (k >> 2)) : __ldg(L4 + (k >> 2));
        const int4 i1 = STREAMED ? __ldcs(L4 + (k >> 2) + 1) : __ldg(L4 + (k >> 2) + 1);
        const int32_t ii[8] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
        double vv[8], xa[8], xb[8];
#pragma unroll
        for (int j = 0; j < 8; j++) vv[j] = STREAMED ? __ldcs(W + nb_w_off(k + j, P)) : __ldg(W + nb_w_off(k + j, P));
#pragma unroll
        for (int j = 0; j < 8; j++) {
            xa[j] = __ldg(x0 + ii[j]);
            if (NRHS == 2) xb[j] = __ldg(x1 + ii[j]);
        }
#pragma unroll
        for (int j = 0; j < 8; j++) {
            a0 += vv[j] * xa[j];
            if (NRHS == 2) a1 += vv[j] * xb[j];
        }
    }
    for (; k + 4 <= K; k += 4) {
        const int4 i0 = STREAMED ? __ldcs(L4 + (k >> 2)) : __ldg(L4 + (k >> 2));
        const int32_t ii[4] = {i0.x, i0.y, i0.z, i0.w};
        double vv[4];
#pragma unroll
        for (int j = 0; j < 4; j++) vv[j] = STREAMED ? __ldcs(W + nb_w_off(k + j, P)) : __ldg(W + nb_w_off(k + j, P));
#pragma unroll
        for (int j = 0; j < 4; j++) {
            a0 += vv[j] * __ldg(x0 + ii[j]);
            if (NRHS == 2) a1 += vv[j] * __ldg(x1 + ii[j]);
        }
    }
    for (; k < K; k++) {
        const int32_t ii = __ldg(L + k);
        const double vv = __ldg(W + nb_w_off(k, P));
        a0 += vv * __ldg(x0 + ii);
        if (NRHS == 2) a1 += vv * __ldg(x1 + ii);
    }
}

template <int NRHS>
__device__ __forceinline__ void nb_row_dot_dict(const StreamArgs& A, int alpha_m1, int64_t row,
                                                const double* __restrict__ x0, const double* __restrict__ x1,
                                                double& y0, double& y1)
{
    const int2 d = __ldcs(A.desc + (int64_t)alpha_m1 * A.desc_stride + row);
    const unsigned cw = (unsigned)d.y;
    const NbDirClass* __restrict__ C = A.cls + alpha_m1 * NB_MAX_CLS + (cw >> NB_PAT_BITS);
    const int K = C->K;
    const double* __restrict__ W = C->W + 2 * (int64_t)(cw & NB_PAT_MASK);
    const int32_t* __restrict__ L = C->L + (int64_t)d.x * C->NL;     // NL = list pitch (K rounded up to 4)
    const int64_t P = C->P;
    double a0 = 0.0, a1 = 0.0;
    if (C->streamed) nb_dict_accumulate<NRHS, true>(W, L, K, P, x0, x1, a0, a1);
    else nb_dict_accumulate<NRHS, false>(W, L, K, P, x0, x1, a0, a1);
    y0 = a0;
    y1 = a1;
}

// ---- cache-policy loads for the staged kernels.  The L1 is shared by three streams of very different reuse:
// weight patterns (re-read by every CTA on the SM: keep), descriptors / staging indices (read once: do not
// allocate), gathered support values (some reuse between neighbouring CTAs: default policy).
__device__ __forceinline__ double nb_ld_keep(const double* p)
{
    double v;
    asm volatile("ld.global.L1::evict_last.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int32_t nb_ld_once(const int32_t* p)
{
    int32_t v;
    asm volatile("ld.global.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int2 nb_ld_once(const int2* p)
{
    int2 v;
    asm volatile("ld.global.L1::no_allocate.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}

// Staged row product: support values from shared memory (xs0/xs1 + off), weights from the pair-packed pattern pool:
// one 16-byte global load feeds two multiply-adds.  Same summation order as the other formats (k = 0..K-1 as
// stored).  All weights of a batch are requested before the first is used; batches are predicated (address clamped),
// so a row of K entries waits ceil(ceil(K/2) / B2) load latencies.
__device__ __forceinline__ double2 nb_ld_keep2(const double2* p)
{
    double2 v;
    asm volatile("ld.global.L1::evict_last.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ double2 nb_ld_stream2(const double2* p)
{
    double2 v;
    asm volatile("ld.global.cs.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

template <int NRHS, bool STREAMED, int B2>
__device__ __forceinline__ void nb_staged_batches(const double2* __restrict__ W2, int K, int64_t P,
                                                  const double* __restrict__ s0, const double* __restrict__ s1,
                                                  double& a0, double& a1)
{
    const int Kh = (K + 1) >> 1;
    for (int kk = 0; kk < Kh; kk += B2) {
        double2 vv[B2];
#pragma unroll
        for (int j = 0; j < B2; j++) {
            const int kj = min(kk + j, Kh - 1);
            vv[j] = STREAMED ? nb_ld_stream2(W2 + (int64_t)kj * P) : nb_ld_keep2(W2 + (int64_t)kj * P);
        }
#pragma unroll
        for (int j = 0; j < B2; j++) {
            const int k = 2 * (kk + j);
            if (k < K) {
                a0 += vv[j].x * s0[k];
                if (NRHS == 2) a1 += vv[j].x * s1[k];
            }
            if (k + 1 < K) {        // the staged lists are not padded: never touch the slot behind an odd list
                a0 += vv[j].y * s0[k + 1];
                if (NRHS == 2) a1 += vv[j].y * s1[k + 1];
            }
        }
    }
}

template <int NRHS, bool STREAMED>
__device__ __forceinline__ void nb_staged_accumulate(const double* __restrict__ W, int K, int64_t P,
                                                     const double* __restrict__ s0, const double* __restrict__ s1,
                                                     double& a0, double& a1)
{
    // K = (p+1)^k on the path: 5 / 25 / 125 for the FE order of the benchmark configurations -> 3 / 13 / 63 pairs
    const double2* W2 = reinterpret_cast<const double2*>(W);
    if (K <= 8) nb_staged_batches<NRHS, STREAMED, 4>(W2, K, P, s0, s1, a0, a1);
    else nb_staged_batches<NRHS, STREAMED, 7>(W2, K, P, s0, s1, a0, a1);
}

template <int NRHS>
__device__ __forceinline__ void nb_row_dot_staged(const StreamArgs& A, int alpha_m1, int2 d, const double* __restrict__ xs0,
                                                  const double* __restrict__ xs1, double& y0, double& y1)
{
    const unsigned dx = (unsigned)d.x;
    const unsigned cls = dx >> 16;
    int K, streamed;
    int64_t P;
    const double* __restrict__ W;
    if (cls == 0) {
        const int kk = A.c0_K[alpha_m1];
        K = kk & 0x3fffffff;
        streamed = kk >> 30;
        P = A.c0_P[alpha_m1];
        W = A.c0_W[alpha_m1] + 2 * (int64_t)(unsigned)d.y;
    } else {
        const NbDirClass* __restrict__ C = A.cls + alpha_m1 * NB_MAX_CLS + cls;
        K = C->K;
        streamed = C->streamed;
        P = C->P;
        W = C->W + 2 * (int64_t)(unsigned)d.y;
    }
    const double* __restrict__ s0 = xs0 + (dx & 0xffffu);
    const double* __restrict__ s1 = xs1 + (dx & 0xffffu);
    double a0 = 0.0, a1 = 0.0;
    if (streamed) nb_staged_accumulate<NRHS, true>(W, K, P, s0, s1, a0, a1);
    else nb_staged_accumulate<NRHS, false>(W, K, P, s0, s1, a0, a1);
    y0 = a0;
    y1 = a1;
}

// Two rows that share class and weight pattern (the same position in two neighbouring cells): every weight is loaded
// once and feeds both rows (and both distributions).  Per row the summation order is unchanged (k = 0..K-1), so the
// results are bit-identical to nb_row_dot_staged.  acc = { row0 f, row1 f, row0 g, row1 g }.
template <int NRHS, bool STREAMED, int B2>
__device__ __forceinline__ void nb_staged_batches_pair(const double2* __restrict__ W2, int K, int64_t P,
                                                       const double* __restrict__ s0, const double* __restrict__ s1,
                                                       const double* __restrict__ g0, const double* __restrict__ g1,
                                                       double (&acc)[4])
{
    const int Kh = (K + 1) >> 1;
    for (int kk = 0; kk < Kh; kk += B2) {
        double2 vv[B2];
#pragma unroll
        for (int j = 0; j < B2; j++) {
            const int kj = min(kk + j, Kh - 1);
            vv[j] = STREAMED ? nb_ld_stream2(W2 + (int64_t)kj * P) : nb_ld_keep2(W2 + (int64_t)kj * P);
        }
#pragma unroll
        for (int j = 0; j < B2; j++) {
            const int k = 2 * (kk + j);
            if (k < K) {
                acc[0] += vv[j].x * s0[k];
                acc[1] += vv[j].x * s1[k];
                if (NRHS == 2) {
                    acc[2] += vv[j].x * g0[k];
                    acc[3] += vv[j].x * g1[k];
                }
            }
            if (k + 1 < K) {
                acc[0] += vv[j].y * s0[k + 1];
                acc[1] += vv[j].y * s1[k + 1];
                if (NRHS == 2) {
                    acc[2] += vv[j].y * g0[k + 1];
                    acc[3] += vv[j].y * g1[k + 1];
                }
            }
        }
    }
}

// Rows r0 and r1 of one direction (descriptors d0, d1) from the staged values xs0 (f) / xs1 (g): shared-weight path when
// both rows have the same class and pattern, two independent products otherwise.  y = { row0 f, row1 f, row0 g, row1 g }.
template <int NRHS>
__device__ __forceinline__ void nb_row_dot_staged_pair(const StreamArgs& A, int alpha_m1, int2 d0, int2 d1,
                                                       const double* __restrict__ xs0, const double* __restrict__ xs1,
                                                       double (&y)[4])
{
    const unsigned x0 = (unsigned)d0.x, x1 = (unsigned)d1.x;
    if ((x0 >> 16) == (x1 >> 16) && d0.y == d1.y) {
        const unsigned cls = x0 >> 16;
        int K, streamed;
        int64_t P;
        const double* __restrict__ W;
        if (cls == 0) {
            const int kk = A.c0_K[alpha_m1];
            K = kk & 0x3fffffff;
            streamed = kk >> 30;
            P = A.c0_P[alpha_m1];
            W = A.c0_W[alpha_m1] + 2 * (int64_t)(unsigned)d0.y;
        } else {
            const NbDirClass* __restrict__ C = A.cls + alpha_m1 * NB_MAX_CLS + cls;
            K = C->K;
            streamed = C->streamed;
            P = C->P;
            W = C->W + 2 * (int64_t)(unsigned)d0.y;
        }
        const double2* W2 = reinterpret_cast<const double2*>(W);
        const unsigned o0 = x0 & 0xffffu, o1 = x1 & 0xffffu;
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        if (streamed) {
            if (K <= 8) nb_staged_batches_pair<NRHS, true, 4>(W2, K, P, xs0 + o0, xs0 + o1, xs1 + o0, xs1 + o1, acc);
            else nb_staged_batches_pair<NRHS, true, 7>(W2, K, P, xs0 + o0, xs0 + o1, xs1 + o0, xs1 + o1, acc);
        } else {
            if (K <= 8) nb_staged_batches_pair<NRHS, false, 4>(W2, K, P, xs0 + o0, xs0 + o1, xs1 + o0, xs1 + o1, acc);
            else nb_staged_batches_pair<NRHS, false, 7>(W2, K, P, xs0 + o0, xs0 + o1, xs1 + o0, xs1 + o1, acc);
        }
#pragma unroll
        for (int i = 0; i < 4; i++) y[i] = acc[i];
    } else {
        nb_row_dot_staged<NRHS>(A, alpha_m1, d0, xs0, xs1, y[0], y[2]);
        nb_row_dot_staged<NRHS>(A, alpha_m1, d1, xs0, xs1, y[1], y[3]);
    }
}

// ---- staging: 8-byte asynchronous copies global -> shared (LDGSTS), so that a thread has all its gathers of a
// pass in flight at once instead of waiting for each batch to land in registers ----
__device__ __forceinline__ void nb_cp_async8(double* smem_dst, const double* gmem_src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void nb_cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Copies the `count` staged values of a pass: xs0[e] = x0[sc[e]] (and xs1[e] = x1[sc[e]]).  A pass holds at most
// NB_STAGE_CAP values = B per thread, so all index loads are in flight together (one latency) and all copies after
// them (one more).
template <int NRHS>
__device__ __forceinline__ void nb_stage_pass(const int32_t* __restrict__ sc, int count, int tid,
                                              const double* __restrict__ x0, const double* __restrict__ x1,
                                              double* __restrict__ xs0, double* __restrict__ xs1)
{
    constexpr int B = NB_STAGE_CAP / NB_CTA_ROWS;
    static_assert(NB_STAGE_CAP_FG <= NB_STAGE_CAP, "one batch must cover a whole pass");
    int32_t idx[B];
#pragma unroll
    for (int j = 0; j < B; j++) idx[j] = (tid + j * NB_CTA_ROWS < count) ? nb_ld_once(sc + tid + j * NB_CTA_ROWS) : -1;
#pragma unroll
    for (int j = 0; j < B; j++)
        if (idx[j] >= 0) {
            nb_cp_async8(xs0 + tid + j * NB_CTA_ROWS, x0 + idx[j]);
            if (NRHS == 2) nb_cp_async8(xs1 + tid + j * NB_CTA_ROWS, x1 + idx[j]);
        }
    nb_cp_async_wait_all();
}


// ---------------------------------------------------------------------------------------------
// NB_FMT_GRID device side: mbarrier + TMA tensor copies, row products with the per-direction offset table
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void nb_mbar_init(uint64_t* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void nb_mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void nb_mbar_expect_tx(uint64_t* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void nb_mbar_wait(uint64_t* bar, unsigned parity)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "NB_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra NB_DONE_%=;\n"
        "bra NB_WAIT_%=;\n"
        "NB_DONE_%=:\n"
        "}\n" ::"r"(a), "r"(parity) : "memory");
}
// box with origin (x, y, z) of the 3-d tensor `tmap` -> shared memory, completion counted in bytes on `bar`
__device__ __forceinline__ void nb_tma_load_3d(double* smem_dst, const void* tmap, int x, int y, int z, uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(tmap), "r"(x), "r"(y), "r"(z),
                   "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

// box of shared memory (dense, x fastest) -> the 3-d tensor `tmap` at origin (x, y, z); points outside the tensor are clipped
__device__ __forceinline__ void nb_tma_store_3d(const void* tmap, const double* smem_src, int x, int y, int z)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(tmap), "r"((unsigned)__cvta_generic_to_shared(smem_src)), "r"(x), "r"(y), "r"(z) : "memory");
}
__device__ __forceinline__ void nb_tma_store_commit_and_wait()
{
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// contiguous bytes global -> shared through the TMA unit (no LSU instructions, no L1 allocation), counted on `bar`;
// addresses and size are multiples of 16
__device__ __forceinline__ void nb_bulk_load(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src), "r"(bytes),
                   "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

// Offsets of a batch of entries: the table holds two 16-bit offsets per word (entries 2j, 2j+1), read from constant
// memory up front with the weights (asm volatile keeps them at the head of the batch, so their latency overlaps the
// weight loads' instead of sitting in front of every shared-memory load).
__device__ __forceinline__ unsigned nb_ldc_u32(const unsigned* p)
{
    unsigned v;
    asm volatile("{\n.reg .u64 t;\ncvta.to.const.u64 t, %1;\nld.const.u32 %0, [t];\n}" : "=r"(v) : "l"(p));
    return v;
}

// support value at byte offset o (an entry of cGridOff) from the row's first value
__device__ __forceinline__ double nb_at(const double* __restrict__ s, unsigned o)
{
    return *reinterpret_cast<const double*>(reinterpret_cast<const char*>(s) + o);
}

// Two class-0 rows with the same weight pattern: one weight load feeds both rows (and both distributions).
// s0/s1: first support value of row 0 / row 1 in the f buffer, g0/g1 in the g buffer.  Summation order k = 0..K-1.
template <int NRHS, bool STREAMED, int B2>
__device__ __forceinline__ void nb_grid_batches_pair(const double2* __restrict__ W2, int K, int64_t P, const unsigned* __restrict__ off2,
                                                     const double* __restrict__ s0, const double* __restrict__ s1,
                                                     const double* __restrict__ g0, const double* __restrict__ g1, double (&acc)[4])
{
    const int Kh = (K + 1) >> 1;
    for (int kk = 0; kk < Kh; kk += B2) {
        double2 vv[B2];
        unsigned oo[B2];
#pragma unroll
        for (int j = 0; j < B2; j++) {
            const int kj = min(kk + j, Kh - 1);
            vv[j] = STREAMED ? nb_ld_stream2(W2 + (int64_t)kj * P) : nb_ld_keep2(W2 + (int64_t)kj * P);
        }
#pragma unroll
        for (int j = 0; j < B2; j++) oo[j] = nb_ldc_u32(off2 + min(kk + j, Kh - 1));
#pragma unroll
        for (int j = 0; j < B2; j++) {
            const int k = 2 * (kk + j);
            if (k < K) {
                const unsigned o = oo[j] & 0xffffu;
                acc[0] += vv[j].x * nb_at(s0, o);
                acc[1] += vv[j].x * nb_at(s1, o);
                if (NRHS == 2) {
                    acc[2] += vv[j].x * nb_at(g0, o);
                    acc[3] += vv[j].x * nb_at(g1, o);
                }
            }
            if (k + 1 < K) {
                const unsigned o = oo[j] >> 16;
                acc[0] += vv[j].y * nb_at(s0, o);
                acc[1] += vv[j].y * nb_at(s1, o);
                if (NRHS == 2) {
                    acc[2] += vv[j].y * nb_at(g0, o);
                    acc[3] += vv[j].y * nb_at(g1, o);
                }
            }
        }
    }
}

// The same with the row length known at compile time (the lengths a tensor-product element produces: (p+1)^m): no clamped
// indices, no per-entry predicates, the weight pointer advanced by additions.  N weight pairs per chunk, the last pair of
// the last chunk half-used when K is odd.  About 4.5 instructions per multiply-add instead of 10.
template <int NRHS, bool STREAMED, int N, bool LAST_HALF>
__device__ __forceinline__ void nb_grid_pair_chunk(const char*& wp, int64_t pitch, const unsigned* __restrict__ off2,
                                                   const double* __restrict__ s0, const double* __restrict__ s1,
                                                   const double* __restrict__ g0, const double* __restrict__ g1, double (&acc)[4])
{
    double2 vv[N];
    unsigned oo[N];
#pragma unroll
    for (int j = 0; j < N; j++) {
        vv[j] = STREAMED ? nb_ld_stream2(reinterpret_cast<const double2*>(wp)) : nb_ld_keep2(reinterpret_cast<const double2*>(wp));
        wp += pitch;
    }
#pragma unroll
    for (int j = 0; j < N; j++) oo[j] = nb_ldc_u32(off2 + j);
#pragma unroll
    for (int j = 0; j < N; j++) {
        {
            const unsigned o = oo[j] & 0xffffu;
            acc[0] += vv[j].x * nb_at(s0, o);
            acc[1] += vv[j].x * nb_at(s1, o);
            if (NRHS == 2) {
                acc[2] += vv[j].x * nb_at(g0, o);
                acc[3] += vv[j].x * nb_at(g1, o);
            }
        }
        if (!(LAST_HALF && j == N - 1)) {
            const unsigned o = oo[j] >> 16;
            acc[0] += vv[j].y * nb_at(s0, o);
            acc[1] += vv[j].y * nb_at(s1, o);
            if (NRHS == 2) {
                acc[2] += vv[j].y * nb_at(g0, o);
                acc[3] += vv[j].y * nb_at(g1, o);
            }
        }
    }
}

template <int NRHS, bool STREAMED, int K>
__device__ __forceinline__ void nb_grid_pair_fixed(const double2* __restrict__ W2, int64_t P, const unsigned* __restrict__ off2,
                                                   const double* __restrict__ s0, const double* __restrict__ s1,
                                                   const double* __restrict__ g0, const double* __restrict__ g1, double (&acc)[4])
{
    constexpr int Kh = (K + 1) / 2;
    constexpr int B2 = Kh <= 8 ? Kh : 7;
    constexpr int NCH = (Kh + B2 - 1) / B2;
    constexpr int LASTN = Kh - (NCH - 1) * B2;
    const char* wp = reinterpret_cast<const char*>(W2);
    const int64_t pitch = P * 16;
    if (NCH > 1) {
#pragma unroll 1
        for (int b = 0; b < NCH - 1; b++) {
            nb_grid_pair_chunk<NRHS, STREAMED, B2, false>(wp, pitch, off2, s0, s1, g0, g1, acc);
            off2 += B2;
        }
    }
    nb_grid_pair_chunk<NRHS, STREAMED, LASTN, (K & 1) != 0>(wp, pitch, off2, s0, s1, g0, g1, acc);
}

template <int NRHS, bool STREAMED, int B2>
__device__ __forceinline__ void nb_grid_batches(const double2* __restrict__ W2, int K, int64_t P, const unsigned* __restrict__ off2,
                                                const double* __restrict__ s0, const double* __restrict__ g0, double& a0, double& a1)
{
    const int Kh = (K + 1) >> 1;
    for (int kk = 0; kk < Kh; kk += B2) {
        double2 vv[B2];
        unsigned oo[B2];
#pragma unroll
        for (int j = 0; j < B2; j++) {
            const int kj = min(kk + j, Kh - 1);
            vv[j] = STREAMED ? nb_ld_stream2(W2 + (int64_t)kj * P) : nb_ld_keep2(W2 + (int64_t)kj * P);
        }
#pragma unroll
        for (int j = 0; j < B2; j++) oo[j] = nb_ldc_u32(off2 + min(kk + j, Kh - 1));
#pragma unroll
        for (int j = 0; j < B2; j++) {
            const int k = 2 * (kk + j);
            if (k < K) {
                const unsigned o = oo[j] & 0xffffu;
                a0 += vv[j].x * nb_at(s0, o);
                if (NRHS == 2) a1 += vv[j].x * nb_at(g0, o);
            }
            if (k + 1 < K) {
                const unsigned o = oo[j] >> 16;
                a0 += vv[j].y * nb_at(s0, o);
                if (NRHS == 2) a1 += vv[j].y * nb_at(g0, o);
            }
        }
    }
}

// One row of direction a from its grid descriptor d: class 0 rows from the staged boxes, "generic" rows (bit 31) from
// their dictionary list in global memory, rows of the K = 0 class give 0.
template <int NRHS>
__device__ __forceinline__ void nb_row_dot_grid(const StreamArgs& A, int a, int2 d, int32_t row, const unsigned* __restrict__ off,
                                                const double* __restrict__ xs0, const double* __restrict__ xs1,
                                                const double* __restrict__ x0, const double* __restrict__ x1, double& y0, double& y1)
{
    const unsigned dx = (unsigned)d.x;
    y0 = y1 = 0.0;
    if (dx >> 31) {
        if (row >= 0) nb_row_dot_dict<NRHS>(A, a, row, x0, x1, y0, y1);
        return;
    }
    if ((dx >> 16) != 0) return;
    const int kk = A.c0_K[a];
    const int K = kk & 0x3fffffff;
    const int64_t P = A.c0_P[a];
    const double2* W2 = reinterpret_cast<const double2*>(A.c0_W[a] + 2 * (int64_t)(unsigned)d.y);
    const unsigned o = dx & 0xffffu;
    if (kk >> 30) {
        if (K <= 8) nb_grid_batches<NRHS, true, 4>(W2, K, P, off, xs0 + o, xs1 + o, y0, y1);
        else nb_grid_batches<NRHS, true, 7>(W2, K, P, off, xs0 + o, xs1 + o, y0, y1);
    } else {
        if (K <= 8) nb_grid_batches<NRHS, false, 4>(W2, K, P, off, xs0 + o, xs1 + o, y0, y1);
        else nb_grid_batches<NRHS, false, 7>(W2, K, P, off, xs0 + o, xs1 + o, y0, y1);
    }
}

// Rows r0 / r1 (descriptors d0 / d1) of direction a; y = { row0 f, row1 f, row0 g, row1 g }.
template <int NRHS>
__device__ __forceinline__ void nb_row_dot_grid_pair(const StreamArgs& A, int a, int2 d0, int2 d1, int32_t r0, int32_t r1,
                                                     const unsigned* __restrict__ off, const double* __restrict__ xs0,
                                                     const double* __restrict__ xs1, const double* __restrict__ x0,
                                                     const double* __restrict__ x1, double (&y)[4])
{
    const unsigned u0 = (unsigned)d0.x, u1 = (unsigned)d1.x;
    if (((u0 | u1) >> 16) == 0 && d0.y == d1.y) {        // both class 0 (not generic), same weight pattern
        const int kk = A.c0_K[a];
        const int K = kk & 0x3fffffff;
        const int64_t P = A.c0_P[a];
        const double2* W2 = reinterpret_cast<const double2*>(A.c0_W[a] + 2 * (int64_t)(unsigned)d0.y);
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        const bool streamed = (kk >> 30) != 0;
#define NB_GRID_FIXED_K(KK)                                                                                                        \
    case KK:                                                                                                                       \
        if (streamed) nb_grid_pair_fixed<NRHS, true, KK>(W2, P, off, xs0 + u0, xs0 + u1, xs1 + u0, xs1 + u1, acc);                 \
        else nb_grid_pair_fixed<NRHS, false, KK>(W2, P, off, xs0 + u0, xs0 + u1, xs1 + u0, xs1 + u1, acc);                        \
        break;
        switch (K) {        // the row lengths of FE orders 2, 3, 4 (one, two, three moving axes)
            NB_GRID_FIXED_K(3) NB_GRID_FIXED_K(4) NB_GRID_FIXED_K(5)
            NB_GRID_FIXED_K(9) NB_GRID_FIXED_K(16) NB_GRID_FIXED_K(25)
#if NB_GRID_K3D
            NB_GRID_FIXED_K(27) NB_GRID_FIXED_K(64) NB_GRID_FIXED_K(125)
#endif
        default:
            if (streamed) {
                if (K <= 8) nb_grid_batches_pair<NRHS, true, 4>(W2, K, P, off, xs0 + u0, xs0 + u1, xs1 + u0, xs1 + u1, acc);
                else nb_grid_batches_pair<NRHS, true, 7>(W2, K, P, off, xs0 + u0, xs0 + u1, xs1 + u0, xs1 + u1, acc);
            } else {
                if (K <= 8) nb_grid_batches_pair<NRHS, false, 4>(W2, K, P, off, xs0 + u0, xs0 + u1, xs1 + u0, xs1 + u1, acc);
                else nb_grid_batches_pair<NRHS, false, 7>(W2, K, P, off, xs0 + u0, xs0 + u1, xs1 + u0, xs1 + u1, acc);
            }
        }
#undef NB_GRID_FIXED_K
#pragma unroll
        for (int i = 0; i < 4; i++) y[i] = acc[i];
    } else {
        nb_row_dot_grid<NRHS>(A, a, d0, r0, off, xs0, xs1, x0, x1, y[0], y[2]);
        nb_row_dot_grid<NRHS>(A, a, d1, r1, off, xs0, xs1, x0, x1, y[1], y[3]);
    }
}

template <int FMT, int NRHS>
__device__ __forceinline__ void nb_row_dot(const StreamArgs& A, int alpha_m1, int64_t row, int64_t slice, int lane,
                                           const double* __restrict__ x0, const double* __restrict__ x1,
                                           double& y0, double& y1)
{
    if (FMT == NB_FMT_ELL) nb_row_dot_ell<NRHS>(A, alpha_m1, slice, lane, x0, x1, y0, y1);
    else nb_row_dot_dict<NRHS>(A, alpha_m1, row, x0, x1, y0, y1);
}
