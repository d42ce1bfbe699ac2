// kernels.cuh -- fused stream+collide and stand-alone collide kernels, templated on the stencil.
// Instantiated once per (D,Q) by inst.cu (one translation unit per stencil so that the
// build parallelises and each unit owns its constant-memory block).
// The production path is k_stream_collide_f_staged / k_stream_collide_fg_staged (staged dictionary format);
// k_stream_collide_f / _fg<FMT> serve the ELL and unstaged dictionary formats (any CSR, also the fallback when the
// rows of a CTA share too little for staging).
#pragma once
#include "stream_common.cuh"
#include "collide.cuh"
#include "entropic.cuh"

// ---------------------------------------------------------------------------------------------
// kernels: hot path
// ---------------------------------------------------------------------------------------------
// Fused stream + collide, f only.  x: current populations (owned + ghosts), y: next.
// Two phases per thread (= one DoF): (1) a compact, not unrolled loop over the Q-1 rows of the DoF that
// parks each streamed value in this thread's column of a shared-memory tile -- few live registers, so many
// warps are resident while the gathers are in flight; (2) the collision on the Q values read back from the
// tile.  No barrier: a thread only ever touches its own column.
#ifndef NB_FUSED_OCC_F
#define NB_FUSED_OCC_F 5
#endif
#ifndef NB_FUSED_OCC_FG
#define NB_FUSED_OCC_FG 2
#endif
template <int D, int Q, int EQ, int FMT>
__global__ void __launch_bounds__(128, NB_FUSED_OCC_F)
k_stream_collide_f(StreamArgs A, const double* __restrict__ x, double* __restrict__ y,
                   double* __restrict__ rho_out, double* __restrict__ u_out, int* __restrict__ flag)
{
    __shared__ double tile[Q][128];
    const int tid = threadIdx.x;
    const int64_t row = blockIdx.x * (int64_t)blockDim.x + tid;
    const int64_t slice = row >> 5;
    const int lane = tid & 31;
    if (slice >= A.n_slices) return;
    const bool active = row < A.n_owned;
    if (FMT == NB_FMT_DICT && !active) return;
    tile[0][tid] = active ? x[row] : 0.0;
#pragma unroll 1
    for (int a = 1; a < Q; a++) {
        double r, dummy;
        nb_row_dot<FMT, 1>(A, a - 1, row, slice, lane, x, nullptr, r, dummy);
        tile[a][tid] = r;
    }
    if (!active) return;
    double f[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) f[q] = tile[q][tid];
    double rho, v[3] = {0.0, 0.0, 0.0};
    if (nb_collide_f<D, Q, EQ>(f, rho, v, false, EQ == NB_KIND_MRT_ENTROPIC ? rho_out[row] : 1.0)) *flag = 1;
#pragma unroll
    for (int q = 0; q < Q; q++) y[(int64_t)q * A.stride + row] = f[q];
    rho_out[row] = rho;
#pragma unroll
    for (int j = 0; j < D; j++) u_out[(int64_t)j * A.n_owned + row] = v[j];
}

// Fused stream + collide for f and g (one pass over the matrix for both distributions).
template <int D, int Q, int EQ, int FMT>
__global__ void __launch_bounds__(128, NB_FUSED_OCC_FG)
k_stream_collide_fg(StreamArgs A, const double* __restrict__ xf, const double* __restrict__ xg,
                    double* __restrict__ yf, double* __restrict__ yg, double* __restrict__ rho_out,
                    double* __restrict__ u_out, double* __restrict__ T_out, double* __restrict__ s_out,
                    int* __restrict__ flag)
{
    extern __shared__ double tile_fg[];        // [2][Q][128]
    double (*tf)[128] = reinterpret_cast<double (*)[128]>(tile_fg);
    double (*tg)[128] = reinterpret_cast<double (*)[128]>(tile_fg + Q * 128);
    const int tid = threadIdx.x;
    const int64_t row = blockIdx.x * (int64_t)blockDim.x + tid;
    const int64_t slice = row >> 5;
    const int lane = tid & 31;
    if (slice >= A.n_slices) return;
    const bool active = row < A.n_owned;
    if (FMT == NB_FMT_DICT && !active) return;
    tf[0][tid] = active ? xf[row] : 0.0;
    tg[0][tid] = active ? xg[row] : 0.0;
#pragma unroll 1
    for (int a = 1; a < Q; a++) {
        double r0, r1;
        nb_row_dot<FMT, 2>(A, a - 1, row, slice, lane, xf, xg, r0, r1);
        tf[a][tid] = r0;
        tg[a][tid] = r1;
    }
    if (!active) return;
    double f[Q], g[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) {
        f[q] = tf[q][tid];
        g[q] = tg[q][tid];
    }
    double rho, u[3], T, sensor;
    nb_collide_bgk_fg<D, Q, EQ>(f, g, rho, u, T, sensor, nullptr);
    if (rho < 1e-10) *flag = 1;
#pragma unroll
    for (int q = 0; q < Q; q++) {
        yf[(int64_t)q * A.stride + row] = f[q];
        yg[(int64_t)q * A.stride + row] = g[q];
    }
    rho_out[row] = rho;
    T_out[row] = T;
    s_out[row] = sensor;
#pragma unroll
    for (int j = 0; j < D; j++) u_out[(int64_t)j * A.n_owned + row] = u[j] * cP.scaling;
}

// ---------------------------------------------------------------------------------------------
// NB_FMT_STAGED variants: the CTA stages the support values of a pass in shared memory, then every thread
// (= one DoF) takes its rows from there.  Dynamic shared memory: [tile(s)][staged values].
// The descriptors of a pass's directions are parked in the tile slots their results will overwrite.
// ---------------------------------------------------------------------------------------------
// 5 CTAs per SM: 102 registers per thread (the BGK epilogue fits without spilling) and 5 x 36 KB of shared memory;
// measured 0.76 ms per step against 0.83 ms at 4 CTAs on configuration 2 (profiles/).
#ifndef NB_STAGED_OCC_F
#define NB_STAGED_OCC_F 5
#endif
__device__ __forceinline__ int2 nb_empty_desc(const StreamArgs& A, int alpha_m1)
{
    // class NB_MAX_CLS-1 of every direction is reserved by the builder as the K = 0 class
    (void)A; (void)alpha_m1;
    return make_int2((int)((unsigned)(NB_MAX_CLS - 1) << 16), 0);
}

template <int D, int Q, int EQ>
__global__ void __launch_bounds__(NB_CTA_ROWS, NB_STAGED_OCC_F)
k_stream_collide_f_staged(StreamArgs A, const double* __restrict__ x, double* __restrict__ y,
                          double* __restrict__ rho_out, double* __restrict__ u_out, int* __restrict__ flag)
{
    extern __shared__ double smem_staged[];
    double (*tile)[NB_CTA_ROWS] = reinterpret_cast<double (*)[NB_CTA_ROWS]>(smem_staged);   // [Q][128]
    double* xs = smem_staged + Q * NB_CTA_ROWS;                                              // [NB_STAGE_CAP]
    const int tid = threadIdx.x;
    const int64_t cta = A.cta_map ? (int64_t)__ldg(A.cta_map + blockIdx.x) : (int64_t)blockIdx.x;
    const int64_t row = cta * NB_CTA_ROWS + tid;
    const bool active = row < A.n_owned;
    tile[0][tid] = active ? x[row] : 0.0;
    const int p0 = __ldg(A.stage_cta + cta), p1 = __ldg(A.stage_cta + cta + 1);
    for (int p = p0; p < p1; p++) {
        const NbStagePass ps = A.stage_pass[p];
        if (p > p0) __syncthreads();          // the previous pass's rows are done with xs
        if (active) {
            const int2* __restrict__ dp = A.sdesc + (int64_t)ps.a0 * A.desc_stride + row;
            for (int a = ps.a0; a < ps.a1; a++, dp += A.desc_stride) nb_cp_async8(&tile[a + 1][tid], reinterpret_cast<const double*>(dp));
        } else {
            for (int a = ps.a0; a < ps.a1; a++) reinterpret_cast<int2*>(&tile[a + 1][tid])[0] = nb_empty_desc(A, a);
        }
        nb_stage_pass<1>(A.stage_col + ps.begin, ps.count, tid, x, x, xs, xs);
        __syncthreads();
        // Rows t and t + 64 of a CTA are the same position in two neighbouring cells whenever the cells are alike, i.e.
        // they use the same weight pattern.  Each half of the CTA therefore takes every other direction for BOTH rows:
        // a weight is loaded once and used twice, which halves the weight traffic through the L1.
        static_assert(NB_CTA_ROWS == 128, "row pairing assumes two 64-row halves");
        const int half = tid >> 6, t0 = tid & 63;
#pragma unroll 1
        for (int a = ps.a0 + ((ps.a0 ^ half) & 1); a < ps.a1; a += 2) {
            const int2 d0 = reinterpret_cast<const int2*>(&tile[a + 1][t0])[0];
            const int2 d1 = reinterpret_cast<const int2*>(&tile[a + 1][t0 + 64])[0];
            double r[4];
            nb_row_dot_staged_pair<1>(A, a, d0, d1, xs, xs, r);
            tile[a + 1][t0] = r[0];
            tile[a + 1][t0 + 64] = r[1];
        }
    }
    __syncthreads();          // results of a row come from the other half of the CTA
    if (!active) return;
    double f[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) f[q] = tile[q][tid];
    double rho, v[3] = {0.0, 0.0, 0.0};
    if (nb_collide_f<D, Q, EQ>(f, rho, v, false, EQ == NB_KIND_MRT_ENTROPIC ? rho_out[row] : 1.0)) *flag = 1;
#pragma unroll
    for (int q = 0; q < Q; q++) y[(int64_t)q * A.stride + row] = f[q];
    rho_out[row] = rho;
#pragma unroll
    for (int j = 0; j < D; j++) u_out[(int64_t)j * A.n_owned + row] = v[j];
}

// f + g: one pass over the matrix for both distributions; [tile f][tile g][xs f][xs g]
template <int D, int Q, int EQ>
__global__ void __launch_bounds__(NB_CTA_ROWS, NB_FUSED_OCC_FG)
k_stream_collide_fg_staged(StreamArgs A, const double* __restrict__ xf, const double* __restrict__ xg,
                           double* __restrict__ yf, double* __restrict__ yg, double* __restrict__ rho_out,
                           double* __restrict__ u_out, double* __restrict__ T_out, double* __restrict__ s_out,
                           int* __restrict__ flag)
{
    extern __shared__ double smem_staged[];
    double (*tf)[NB_CTA_ROWS] = reinterpret_cast<double (*)[NB_CTA_ROWS]>(smem_staged);
    double (*tg)[NB_CTA_ROWS] = reinterpret_cast<double (*)[NB_CTA_ROWS]>(smem_staged + Q * NB_CTA_ROWS);
    double* xsf = smem_staged + 2 * Q * NB_CTA_ROWS;
    double* xsg = xsf + NB_STAGE_CAP_FG;
    const int tid = threadIdx.x;
    const int64_t cta = A.cta_map ? (int64_t)__ldg(A.cta_map + blockIdx.x) : (int64_t)blockIdx.x;
    const int64_t row = cta * NB_CTA_ROWS + tid;
    const bool active = row < A.n_owned;
    tf[0][tid] = active ? xf[row] : 0.0;
    tg[0][tid] = active ? xg[row] : 0.0;
    const int p0 = __ldg(A.stage_cta + cta), p1 = __ldg(A.stage_cta + cta + 1);
    for (int p = p0; p < p1; p++) {
        const NbStagePass ps = A.stage_pass[p];
        if (p > p0) __syncthreads();
        if (active) {
            const int2* __restrict__ dp = A.sdesc + (int64_t)ps.a0 * A.desc_stride + row;
            for (int a = ps.a0; a < ps.a1; a++, dp += A.desc_stride) nb_cp_async8(&tf[a + 1][tid], reinterpret_cast<const double*>(dp));
        } else {
            for (int a = ps.a0; a < ps.a1; a++) reinterpret_cast<int2*>(&tf[a + 1][tid])[0] = nb_empty_desc(A, a);
        }
        nb_stage_pass<2>(A.stage_col + ps.begin, ps.count, tid, xf, xg, xsf, xsg);
        __syncthreads();
        // row pairing as in k_stream_collide_f_staged: one weight load feeds two rows and two distributions
        const int half = tid >> 6, t0 = tid & 63;
#pragma unroll 1
        for (int a = ps.a0 + ((ps.a0 ^ half) & 1); a < ps.a1; a += 2) {
            const int2 d0 = reinterpret_cast<const int2*>(&tf[a + 1][t0])[0];
            const int2 d1 = reinterpret_cast<const int2*>(&tf[a + 1][t0 + 64])[0];
            double r[4];
            nb_row_dot_staged_pair<2>(A, a, d0, d1, xsf, xsg, r);
            tf[a + 1][t0] = r[0];
            tf[a + 1][t0 + 64] = r[1];
            tg[a + 1][t0] = r[2];
            tg[a + 1][t0 + 64] = r[3];
        }
    }
    __syncthreads();          // results of a row come from the other half of the CTA
    if (!active) return;
    double f[Q], g[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) {
        f[q] = tf[q][tid];
        g[q] = tg[q][tid];
    }
    double rho, u[3], T, sensor;
    nb_collide_bgk_fg<D, Q, EQ>(f, g, rho, u, T, sensor, nullptr);
    if (rho < 1e-10) *flag = 1;
#pragma unroll
    for (int q = 0; q < Q; q++) {
        yf[(int64_t)q * A.stride + row] = f[q];
        yg[(int64_t)q * A.stride + row] = g[q];
    }
    rho_out[row] = rho;
    T_out[row] = T;
    s_out[row] = sensor;
#pragma unroll
    for (int j = 0; j < D; j++) u_out[(int64_t)j * A.n_owned + row] = u[j] * cP.scaling;
}


// ---------------------------------------------------------------------------------------------
// NB_FMT_GRID variants (nb200_set_dof_grid): one CTA = one tile of grid points (two half-tiles of whole cells that are
// neighbours in x).  Per pass one thread arms an mbarrier with the pass's byte count and issues the TMA tensor copies
// (boxes of the lexicographic grid copy of the populations) of the NEXT pass into the other staging buffer, then all
// threads wait for the current pass's barrier and take their rows from shared memory: a class-0 row's k-th support value
// sits at (row offset + cGridOff[direction][k]).  Results are written to the canonical arrays and to the grid copy that
// the next step's TMA reads.  Dynamic shared memory: [tile(s)][2 staging buffers (per distribution)].
// ---------------------------------------------------------------------------------------------
// two 16-bit BYTE offsets per word: entries 2j (low half) and 2j+1 (high half) of a class-0 row of the direction
__constant__ unsigned cGridOff[NB_MAX_DIRS][NB_GRID_MAXK / 2];

#ifndef NB_GRID_OCC_F
#define NB_GRID_OCC_F 5
#endif
#ifndef NB_GRID_OCC_FG
#define NB_GRID_OCC_FG 3
#endif

// Issues the TMA copies of pass ps into the staging buffer(s) and arms the barrier.  Called by ALL lanes of one warp:
// lane 0 arms the barrier with the pass's byte count, lane l fetches box l (l + 32, ...) and issues its copy, so the
// table reads of the boxes are in flight together instead of one load latency per box.
template <int NRHS>
__device__ __forceinline__ void nb_grid_issue(const StreamArgs& A, const NbGridPass& ps, double* xs0, double* xs1, uint64_t* bar, int lane)
{
    if (lane == 0) nb_mbar_expect_tx(bar, (unsigned)ps.bytes * NRHS);
    for (int b = lane; b < ps.n_box; b += 32) {
        const NbGridBox bx = A.gbox[ps.box_begin + b];
        nb_tma_load_3d(xs0 + bx.smem_off, reinterpret_cast<const char*>(A.tmap_f) + 128 * (int)bx.dir, bx.x, bx.y, bx.z, bar);
        if (NRHS == 2) nb_tma_load_3d(xs1 + bx.smem_off, reinterpret_cast<const char*>(A.tmap_g) + 128 * (int)bx.dir, bx.x, bx.y, bx.z, bar);
    }
}

// A warp is done with staging buffer `buf` of pass p: the last of the CTA's warps to say so issues the copies of pass
// p + 2 into it (nobody waits for anybody: fast warps run ahead into the other buffer).  cnt: per-buffer arrival counter.
template <int NRHS>
__device__ __forceinline__ void nb_grid_release(const StreamArgs& A, int p, int p1, int buf, int lane, int* cnt, double* xs0, double* xs1,
                                                int cap, uint64_t* mbar)
{
    __syncwarp();
    int last = 0;
    if (lane == 0) {
        __threadfence_block();                      // this warp's reads of the buffer are done before the count says so
        const int old = atomicAdd(&cnt[buf], 1);
        if (old == NB_CTA_ROWS / 32 - 1) {
            cnt[buf] = 0;
            last = 1;
        }
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (last && p + 2 < p1) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy reads before the async-proxy writes
        nb_grid_issue<NRHS>(A, A.gpass[p + 2], xs0 + buf * cap, xs1 + buf * cap, &mbar[buf], lane);
    }
}

template <int D, int Q, int EQ>
__global__ void __launch_bounds__(NB_CTA_ROWS, NB_GRID_OCC_F)
k_stream_collide_f_grid(StreamArgs A, const double* __restrict__ x, double* __restrict__ y, double* __restrict__ ygrid,
                        double* __restrict__ rho_out, double* __restrict__ u_out, int* __restrict__ flag)
{
    extern __shared__ __align__(128) double smem_grid[];
    __shared__ uint64_t mbar[3];
    __shared__ int cnt[2];
    __shared__ int s_tflags;
    __shared__ int32_t srow[NB_CTA_ROWS];
    double (*tile)[NB_CTA_ROWS] = reinterpret_cast<double (*)[NB_CTA_ROWS]>(smem_grid);    // [Q][128]
    double* xs = smem_grid + Q * NB_CTA_ROWS;                                               // [2][A.grid_cap]
    const int tid = threadIdx.x;
    const int64_t tl = A.cta_map ? (int64_t)__ldg(A.cta_map + blockIdx.x) : (int64_t)blockIdx.x;
    const int64_t slot = tl * NB_CTA_ROWS + tid;
    const int32_t row = __ldg(A.tile_row + slot);
    const bool active = row >= 0;
    srow[tid] = row;
    tile[0][tid] = active ? x[row] : 0.0;
    if (tid == 0) {
        nb_mbar_init(&mbar[0], 1);
        nb_mbar_init(&mbar[1], 1);
        nb_mbar_init(&mbar[2], 1);
        cnt[0] = cnt[1] = 0;
        s_tflags = A.tile_store ? (int)__ldg(reinterpret_cast<const short*>(A.tile_store) + tl * 4 + 3) : 0;
        nb_mbar_fence_init();
    }
    __syncthreads();
    const int p0 = __ldg(A.stage_cta + tl), p1 = __ldg(A.stage_cta + tl + 1);
    if (tid < 32) {      // both buffers are free: the first two passes start right away (warp 0 issues the copies)
        if (p0 < p1) nb_grid_issue<1>(A, A.gpass[p0], xs, xs, &mbar[0], tid);
        if (p0 + 1 < p1) nb_grid_issue<1>(A, A.gpass[p0 + 1], xs + A.grid_cap, xs, &mbar[1], tid);
    } else if (tid < 64) {
        // the descriptors of all directions, one bulk copy of 128 x 8 bytes per direction: each parks in the tile row its
        // results will overwrite (no LSU instructions and no L1 allocation: the L1 is left to the weight patterns)
        const int lane = tid - 32;
        if (lane == 0) nb_mbar_expect_tx(&mbar[2], (unsigned)((Q - 1) * NB_CTA_ROWS * sizeof(int2)));
        __syncwarp();
        for (int a = lane; a < Q - 1; a += 32)
            nb_bulk_load(&tile[a + 1][0], A.sdesc + (int64_t)a * A.gdesc_stride + tl * NB_CTA_ROWS, NB_CTA_ROWS * sizeof(int2), &mbar[2]);
    }
    nb_mbar_wait(&mbar[2], 0u);          // descriptors (also those of the partner rows) are in place
    const int half = tid >> 6, t0 = tid & 63;
    const int32_t r0 = srow[t0], r1 = srow[t0 + 64];
    for (int p = p0; p < p1; p++) {
        const int buf = (p - p0) & 1;
        const NbGridPass ps = A.gpass[p];
        nb_mbar_wait(&mbar[buf], (unsigned)(((p - p0) >> 1) & 1));
        const double* __restrict__ xb = xs + buf * A.grid_cap;
#pragma unroll 1
        for (int a = ps.a0 + ((ps.a0 ^ half) & 1); a < ps.a1; a += 2) {
            const int2 d0 = reinterpret_cast<const int2*>(&tile[a + 1][t0])[0];
            const int2 d1 = reinterpret_cast<const int2*>(&tile[a + 1][t0 + 64])[0];
            double r[4];
            nb_row_dot_grid_pair<1>(A, a, d0, d1, r0, r1, cGridOff[a], xb, xb, x, x, r);
            tile[a + 1][t0] = r[0];
            tile[a + 1][t0 + 64] = r[1];
        }
        nb_grid_release<1>(A, p, p1, buf, tid & 31, cnt, xs, xs, A.grid_cap, mbar);
    }
    __syncthreads();          // results of a row come from the other half of the CTA
    // a half-tile of whole cells is a box of the grid copy: where the builder allows it the half goes out as ONE TMA store per
    // population from the tile in shared memory instead of 32-byte pieces from every thread (6 store wavefronts per warp and
    // population on the L1 data pipe); the canonical array is written by the threads either way (contiguous rows)
    if (active) {
        double f[Q];
#pragma unroll
        for (int q = 0; q < Q; q++) f[q] = tile[q][tid];
        double rho, v[3] = {0.0, 0.0, 0.0};
        if (nb_collide_f<D, Q, EQ>(f, rho, v, false, EQ == NB_KIND_MRT_ENTROPIC ? rho_out[row] : 1.0)) *flag = 1;
        // second copy: into the tile (box store below) or straight into the grid copy
        double* __restrict__ y2 = ((s_tflags >> half) & 1) ? &tile[0][tid] : ygrid + __ldg(A.tile_gidx + slot);
        const int64_t pitch2 = ((s_tflags >> half) & 1) ? (int64_t)NB_CTA_ROWS : A.gstride;
#pragma unroll
        for (int q = 0; q < Q; q++) {
            y[(int64_t)q * A.stride + row] = f[q];
            y2[(int64_t)q * pitch2] = f[q];
        }
        rho_out[row] = rho;
#pragma unroll
        for (int j = 0; j < D; j++) u_out[(int64_t)j * A.n_owned + row] = v[j];
    }
    if (s_tflags) {           // uniform over the CTA
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the tile writes above, before the async-proxy reads
        __syncthreads();
        if (t0 == 0 && ((s_tflags >> half) & 1)) {
            const short4 ts = __ldg(A.tile_store + tl);
            const int bx = (int)ts.x + half * A.half_x;
#pragma unroll 1
            for (int q = 0; q < Q; q++)
                nb_tma_store_3d(reinterpret_cast<const char*>(A.tmap_out_f) + 128 * q, &tile[q][half * (NB_CTA_ROWS / 2)], bx, (int)ts.y, (int)ts.z);
            nb_tma_store_commit_and_wait();
        }
    }
}

// f + g: both distributions staged by the same boxes; [tile f][tile g][2 buffers f][2 buffers g]
template <int D, int Q, int EQ>
__global__ void __launch_bounds__(NB_CTA_ROWS, NB_GRID_OCC_FG)
k_stream_collide_fg_grid(StreamArgs A, const double* __restrict__ xf, const double* __restrict__ xg,
                         double* __restrict__ yf, double* __restrict__ yg, double* __restrict__ yfgrid, double* __restrict__ yggrid,
                         double* __restrict__ rho_out, double* __restrict__ u_out, double* __restrict__ T_out,
                         double* __restrict__ s_out, int* __restrict__ flag)
{
    extern __shared__ __align__(128) double smem_grid[];
    __shared__ uint64_t mbar[3];
    __shared__ int cnt[2];
    __shared__ int32_t srow[NB_CTA_ROWS];
    __shared__ double ctab[Q * NB_CT_PITCH];
    nb_ct_fill<D, Q>(ctab, threadIdx.x, NB_CTA_ROWS);       // visible after the barriers below
    double (*tf)[NB_CTA_ROWS] = reinterpret_cast<double (*)[NB_CTA_ROWS]>(smem_grid);
    double (*tg)[NB_CTA_ROWS] = reinterpret_cast<double (*)[NB_CTA_ROWS]>(smem_grid + Q * NB_CTA_ROWS);
    double* xsf = smem_grid + 2 * Q * NB_CTA_ROWS;       // [2][NB_GRID_CAP_FGF]
    double* xsg = xsf + 2 * NB_GRID_CAP_FGF;
    const int tid = threadIdx.x;
    const int64_t tl = A.cta_map ? (int64_t)__ldg(A.cta_map + blockIdx.x) : (int64_t)blockIdx.x;
    const int64_t slot = tl * NB_CTA_ROWS + tid;
    const int32_t row = __ldg(A.tile_row + slot);
    const bool active = row >= 0;
    srow[tid] = row;
    tf[0][tid] = active ? xf[row] : 0.0;
    tg[0][tid] = active ? xg[row] : 0.0;
    if (tid == 0) {
        nb_mbar_init(&mbar[0], 1);
        nb_mbar_init(&mbar[1], 1);
        nb_mbar_init(&mbar[2], 1);
        cnt[0] = cnt[1] = 0;
        nb_mbar_fence_init();
    }
    __syncthreads();
    const int p0 = __ldg(A.stage_cta + tl), p1 = __ldg(A.stage_cta + tl + 1);
    if (tid < 32) {
        if (p0 < p1) nb_grid_issue<2>(A, A.gpass[p0], xsf, xsg, &mbar[0], tid);
        if (p0 + 1 < p1) nb_grid_issue<2>(A, A.gpass[p0 + 1], xsf + NB_GRID_CAP_FGF, xsg + NB_GRID_CAP_FGF, &mbar[1], tid);
    } else if (tid < 64) {      // descriptors by bulk copies, as in k_stream_collide_f_grid
        const int lane = tid - 32;
        if (lane == 0) nb_mbar_expect_tx(&mbar[2], (unsigned)((Q - 1) * NB_CTA_ROWS * sizeof(int2)));
        __syncwarp();
        for (int a = lane; a < Q - 1; a += 32)
            nb_bulk_load(&tf[a + 1][0], A.sdesc + (int64_t)a * A.gdesc_stride + tl * NB_CTA_ROWS, NB_CTA_ROWS * sizeof(int2), &mbar[2]);
    }
    nb_mbar_wait(&mbar[2], 0u);
    const int half = tid >> 6, t0 = tid & 63;
    const int32_t r0 = srow[t0], r1 = srow[t0 + 64];
    for (int p = p0; p < p1; p++) {
        const int buf = (p - p0) & 1;
        const NbGridPass ps = A.gpass[p];
        nb_mbar_wait(&mbar[buf], (unsigned)(((p - p0) >> 1) & 1));
        const double* __restrict__ xbf = xsf + buf * NB_GRID_CAP_FGF;
        const double* __restrict__ xbg = xsg + buf * NB_GRID_CAP_FGF;
#pragma unroll 1
        for (int a = ps.a0 + ((ps.a0 ^ half) & 1); a < ps.a1; a += 2) {
            const int2 d0 = reinterpret_cast<const int2*>(&tf[a + 1][t0])[0];
            const int2 d1 = reinterpret_cast<const int2*>(&tf[a + 1][t0 + 64])[0];
            double r[4];
            nb_row_dot_grid_pair<2>(A, a, d0, d1, r0, r1, cGridOff[a], xbf, xbg, xf, xg, r);
            tf[a + 1][t0] = r[0];
            tf[a + 1][t0 + 64] = r[1];
            tg[a + 1][t0] = r[2];
            tg[a + 1][t0 + 64] = r[3];
        }
        nb_grid_release<2>(A, p, p1, buf, tid & 31, cnt, xsf, xsg, NB_GRID_CAP_FGF, mbar);
    }
    __syncthreads();
    if (!active) return;
    // collision straight from the tiles (collide.cuh: nb_collide_fg_loops): no population arrays in registers
    double rho, u[3], T, sensor;
    const int64_t gi = __ldg(A.tile_gidx + slot);
    auto ldf = [&](int i) { return tf[i][tid]; };
    auto ldg = [&](int i) { return tg[i][tid]; };
    auto st = [&](int i, double fv, double gv) {
        yf[(int64_t)i * A.stride + row] = fv;
        yg[(int64_t)i * A.stride + row] = gv;
        yfgrid[(int64_t)i * A.gstride + gi] = fv;
        yggrid[(int64_t)i * A.gstride + gi] = gv;
    };
    auto stf = [&](int, double) {};
    nb_collide_fg_loops<D, Q, EQ, false>(ctab, ldf, ldg, st, stf, rho, u, T, sensor, nullptr, nullptr);
    if (rho < 1e-10) *flag = 1;
    rho_out[row] = rho;
    T_out[row] = T;
    s_out[row] = sensor;
#pragma unroll
    for (int j = 0; j < D; j++) u_out[(int64_t)j * A.n_owned + row] = u[j] * cP.scaling;
}

// Stream only over the grid copy (the vmult site and the unfused configurations): canonical output only.
// (templated on the stencil so that every stencil unit launches its own instance, which reads that unit's cGridOff)
// (4 CTAs/SM as the register target: left to itself ptxas settles at ~70 registers and issues the weight loads of a batch one
// by one between the multiply-adds instead of all up front)
template <int D, int Q, int NRHS>
__global__ void __launch_bounds__(NB_CTA_ROWS, 4)
k_stream_grid(StreamArgs A, const double* __restrict__ x0, const double* __restrict__ x1,
              double* __restrict__ y0, double* __restrict__ y1)
{
    extern __shared__ __align__(128) double smem_grid[];
    __shared__ uint64_t mbar[2];
    __shared__ int cnt[2];
    __shared__ int32_t srow[NB_CTA_ROWS];
    constexpr int CAP = NB_GRID_CAP_OF(Q, NRHS);
    double* xs0 = smem_grid;                         // [2][CAP]
    double* xs1 = smem_grid + (NRHS == 2 ? 2 * CAP : 0);
    const int tid = threadIdx.x;
    const int64_t tl = A.cta_map ? (int64_t)__ldg(A.cta_map + blockIdx.x) : (int64_t)blockIdx.x;
    const int64_t slot = tl * NB_CTA_ROWS + tid;
    const int32_t row = __ldg(A.tile_row + slot);
    srow[tid] = row;
    if (row >= 0) {
        y0[row] = x0[row];
        if (NRHS == 2) y1[row] = x1[row];
    }
    if (tid == 0) {
        nb_mbar_init(&mbar[0], 1);
        nb_mbar_init(&mbar[1], 1);
        cnt[0] = cnt[1] = 0;
        nb_mbar_fence_init();
    }
    __syncthreads();
    const int p0 = __ldg(A.stage_cta + tl), p1 = __ldg(A.stage_cta + tl + 1);
    if (tid < 32) {
        if (p0 < p1) nb_grid_issue<NRHS>(A, A.gpass[p0], xs0, xs1, &mbar[0], tid);
        if (p0 + 1 < p1) nb_grid_issue<NRHS>(A, A.gpass[p0 + 1], xs0 + CAP, xs1 + CAP, &mbar[1], tid);
    }
    const int half = tid >> 6, t0 = tid & 63;
    const int32_t r0 = srow[t0], r1 = srow[t0 + 64];
    for (int p = p0; p < p1; p++) {
        const int buf = (p - p0) & 1;
        const NbGridPass ps = A.gpass[p];
        nb_mbar_wait(&mbar[buf], (unsigned)(((p - p0) >> 1) & 1));
        const double* __restrict__ xb0 = xs0 + buf * CAP;
        const double* __restrict__ xb1 = xs1 + buf * CAP;
#pragma unroll 1
        for (int a = ps.a0 + ((ps.a0 ^ half) & 1); a < ps.a1; a += 2) {
            const int2 d0 = nb_ld_once(A.sdesc + (int64_t)a * A.gdesc_stride + tl * NB_CTA_ROWS + t0);
            const int2 d1 = nb_ld_once(A.sdesc + (int64_t)a * A.gdesc_stride + tl * NB_CTA_ROWS + t0 + 64);
            double r[4];
            nb_row_dot_grid_pair<NRHS>(A, a, d0, d1, r0, r1, cGridOff[a], xb0, xb1, x0, x1, r);
            if (r0 >= 0) {
                y0[(int64_t)(a + 1) * A.stride + r0] = r[0];
                if (NRHS == 2) y1[(int64_t)(a + 1) * A.stride + r0] = r[2];
            }
            if (r1 >= 0) {
                y0[(int64_t)(a + 1) * A.stride + r1] = r[1];
                if (NRHS == 2) y1[(int64_t)(a + 1) * A.stride + r1] = r[3];
            }
        }
        nb_grid_release<NRHS>(A, p, p1, buf, tid & 31, cnt, xs0, xs1, CAP, mbar);
    }
}

// Stand-alone collide (in place), f only.  FORCE: external-force hooks compiled in.
template <int D, int Q, int EQ, bool FORCE>
__global__ void __launch_bounds__(128)
k_collide_f(int64_t n, int64_t stride, double* __restrict__ fbuf, double* __restrict__ rho_out,
            double* __restrict__ u_out, int in_init, int* __restrict__ flag)
{
    const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (row >= n) return;
    double f[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) f[q] = fbuf[(int64_t)q * stride + row];
    double rho, v[3] = {0.0, 0.0, 0.0};
    if (in_init) {
#pragma unroll
        for (int j = 0; j < D; j++) v[j] = u_out[(int64_t)j * n + row];
    }
    bool bad;
    if constexpr (FORCE) bad = nb_collide_f_forced<D, Q, EQ>(f, rho, v, in_init != 0);
    else bad = nb_collide_f<D, Q, EQ>(f, rho, v, in_init != 0, EQ == NB_KIND_MRT_ENTROPIC ? rho_out[row] : 1.0);
    if (bad) *flag = 1;
#pragma unroll
    for (int q = 0; q < Q; q++) fbuf[(int64_t)q * stride + row] = f[q];
    rho_out[row] = rho;
    // the exact-difference force shifts the stored velocity even in the initialization procedure
    // (postCollisionApplyForces runs unconditionally, CollisionOperator.h:93-96)
    if (!in_init || (FORCE && cP.force_type == 2)) {
#pragma unroll
        for (int j = 0; j < D; j++) u_out[(int64_t)j * n + row] = v[j];
    }
}

// Stand-alone collide (in place), f and g.  gidx != null: the new populations also go into the grid copies (fgrid / ggrid,
// pitch gstride) the TMA kernels read, so that no separate canonical -> grid copy is needed before the next stream.
// Q > 25 takes the looping form (collide.cuh: nb_collide_fg_loops), which keeps no population array in registers.
template <int D, int Q, int EQ, bool FORCE>
__global__ void __launch_bounds__(128)
k_collide_fg(int64_t n, int64_t stride, double* fbuf, double* gbuf,
             double* __restrict__ rho_out, double* __restrict__ u_out, double* __restrict__ T_out,
             double* __restrict__ s_out, int in_init, int* __restrict__ flag,
             const int32_t* __restrict__ gidx, double* __restrict__ fgrid, double* __restrict__ ggrid, int64_t gstride)
{
    const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    __shared__ double ctab[Q > 25 ? Q * NB_CT_PITCH : 1];
    if constexpr (Q > 25) {
        nb_ct_fill<D, Q>(ctab, threadIdx.x, blockDim.x);
        __syncthreads();
    }
    if (row >= n) return;
    double rho, u[3], uo[3], T, sensor, vf[3] = {0.0, 0.0, 0.0};
    if (in_init) {
#pragma unroll
        for (int j = 0; j < D; j++) uo[j] = u_out[(int64_t)j * n + row];
    }
    const int64_t gi = gidx ? (int64_t)gidx[row] : 0;
    if constexpr (Q > 25) {
        double* fp = fbuf + row;
        double* gp = gbuf + row;
        auto ldf = [&](int i) { return fp[(int64_t)i * stride]; };
        auto ldg = [&](int i) { return gp[(int64_t)i * stride]; };
        auto st = [&](int i, double fv, double gv) {
            fp[(int64_t)i * stride] = fv;
            gp[(int64_t)i * stride] = gv;
            if (gidx && !(FORCE && cP.force_type == 2)) fgrid[(int64_t)i * gstride + gi] = fv;
            if (gidx) ggrid[(int64_t)i * gstride + gi] = gv;
        };
        auto stf = [&](int i, double df) {
            const double fv = fp[(int64_t)i * stride] + df;
            fp[(int64_t)i * stride] = fv;
            if (gidx) fgrid[(int64_t)i * gstride + gi] = fv;
        };
        nb_collide_fg_loops<D, Q, EQ, FORCE>(ctab, ldf, ldg, st, stf, rho, u, T, sensor, in_init ? uo : nullptr, vf);
    } else {
        double f[Q], g[Q];
#pragma unroll
        for (int q = 0; q < Q; q++) {
            f[q] = fbuf[(int64_t)q * stride + row];
            g[q] = gbuf[(int64_t)q * stride + row];
        }
        nb_collide_bgk_fg<D, Q, EQ, FORCE>(f, g, rho, u, T, sensor, in_init ? uo : nullptr, vf);
#pragma unroll
        for (int q = 0; q < Q; q++) {
            fbuf[(int64_t)q * stride + row] = f[q];
            gbuf[(int64_t)q * stride + row] = g[q];
            if (gidx) {
                fgrid[(int64_t)q * gstride + gi] = f[q];
                ggrid[(int64_t)q * gstride + gi] = g[q];
            }
        }
    }
    if (rho < 1e-10) *flag = 1;
    rho_out[row] = rho;
    T_out[row] = T;
    s_out[row] = sensor;
    if (FORCE) {
        // u carries the equilibrium shift; the global vector gets raw moment * scaling + 0.5 dt F / rho (vf)
        if (!in_init || cP.force_type == 2) {
#pragma unroll
            for (int j = 0; j < D; j++) u_out[(int64_t)j * n + row] = vf[j];
        }
    } else if (!in_init) {
#pragma unroll
        for (int j = 0; j < D; j++) u_out[(int64_t)j * n + row] = u[j] * cP.scaling;
    }
}

// Wall hits after streaming (SemiLagrangianBoundaryHandler::operate, L/boundaries/SemiLagrangianBoundaryHandler.cpp:43-109).
// Both wall models on the path only touch the destination DoF, so one thread owns one DoF and replays that DoF's
// hits in the reference's iteration order; different DoFs are independent.
//   kind 0  VelocityNeqBounceBack (VelocityNeqBounceBack.cpp:137-195): f[dir](idx) += value (= 2 w rho e.u_wall / cs2,
//           evaluated by the host that owns the wall-velocity function)
//   kind 1  ThermalBounceBack (ThermalBounceBack.cpp:50-109, D3Q45, gamma = 1.4 hard-wired there): f and g of the DoF
//           are re-equilibrated to the wall temperature `value`
template <int D, int Q>
__global__ void __launch_bounds__(64)
k_wall_hits(int64_t n_groups, const int32_t* __restrict__ group_dof, const int64_t* __restrict__ group_off,
            const int32_t* __restrict__ hit_dir, const int32_t* __restrict__ hit_kind, const double* __restrict__ hit_val,
            int64_t stride, double* __restrict__ fbuf, double* __restrict__ gbuf,
            const int32_t* __restrict__ gidx, double* __restrict__ ggrid, int64_t gstride)
{
    const int64_t grp = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (grp >= n_groups) return;
    const int64_t idx = group_dof[grp];
    for (int64_t h = group_off[grp]; h < group_off[grp + 1]; h++) {
        const int kind = hit_kind[h];
        if (kind == 0) {
            double* p = fbuf + (int64_t)hit_dir[h] * stride + idx;
            *p = *p + hit_val[h];
        } else if constexpr (D == 3 && Q == 45) {
            const double gamma = 1.4, Tw = hit_val[h];
            double fd[Q], gd[Q], feq[Q];
#pragma unroll
            for (int i = 0; i < Q; i++) {
                fd[i] = fbuf[(int64_t)i * stride + idx];
                gd[i] = gbuf[(int64_t)i * stride + idx];
            }
            const double rho = nb_density<Q>(fd);
            double u[3] = {0.0, 0.0, 0.0};
#pragma unroll
            for (int j = 0; j < D; j++) {
                double sacc = 0.0;
#pragma unroll
                for (int i = 0; i < Q; i++) sacc += cP.e[i][j] * fd[i];
                u[j] = sacc * 1.0 / rho;
            }
            double T = 0.0;
#pragma unroll
            for (int i = 0; i < Q; i++) {
                double sum = 0.0;
#pragma unroll
                for (int a = 0; a < D; a++) sum += (cP.e[i][a] - u[a]) * (cP.e[i][a] - u[a]);
                T += sum * fd[i] * cP.inv_cs2 + gd[i];
            }
            const double C_v = 1. / (gamma - 1.0);
            T = T * 0.5 / (rho * C_v);
            if (fabs(T - Tw) > 0.00001) {
                nb_feq_quartic<D, Q>(rho, u, T, feq);
#pragma unroll
                for (int i = 0; i < Q; i++) fd[i] -= feq[i];
                nb_feq_quartic<D, Q>(rho, u, Tw, feq);
                const double gfac = (Tw) * (2.0 * C_v - D);
#pragma unroll
                for (int i = 0; i < Q; i++) {
                    fbuf[(int64_t)i * stride + idx] = fd[i] + feq[i];
                    gbuf[(int64_t)i * stride + idx] = feq[i] * gfac;
                    if (gidx) ggrid[(int64_t)i * gstride + gidx[idx]] = feq[i] * gfac;
                }
            }
        }
    }
}

// DataProcessor hook after collide (CFDSolver.cpp:889-891): f <- A f per owned DoF, A in constant memory
// (PseudoEntropicStabilizer::apply_d2q9 / apply_d2q9_with_e / apply_d3q19, PseudoEntropicStabilizer.cpp:152-262).
// Row sums in the reference's order: j = 0..Q-1 accumulated from 0.
template <int Q>
__global__ void __launch_bounds__(128)
k_post_matrix(int64_t n, int64_t stride, double* __restrict__ fbuf)
{
    const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (row >= n) return;
    if constexpr (Q <= NB_MRT_MAXQ) {
        double f[Q];
#pragma unroll
        for (int q = 0; q < Q; q++) f[q] = fbuf[(int64_t)q * stride + row];
#pragma unroll
        for (int i = 0; i < Q; i++) {
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < Q; j++) acc += cA[i][j] * f[j];
            fbuf[(int64_t)i * stride + row] = acc;
        }
    }
}

// Conserved sums: deterministic two-stage reduction.  partial[blk*5 + m].
template <int D, int Q>
__global__ void __launch_bounds__(256)
k_conserved_partial(int64_t n, int64_t stride, const double* __restrict__ f,
                    const double* __restrict__ g, double* __restrict__ partial)
{
    __shared__ double sm[5][256];
    double acc[5] = {0, 0, 0, 0, 0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double rho = 0.0, m[3] = {0, 0, 0}, e2 = 0.0, gs = 0.0;
        for (int q = 0; q < Q; q++) {
            const double v = f[(int64_t)q * stride + i];
            rho += v;
            double ee = 0.0;
            for (int j = 0; j < D; j++) {
                m[j] += cP.es[q][j] * v;
                ee += cP.e[q][j] * cP.e[q][j];
            }
            e2 += ee * v;
            if (g) gs += g[(int64_t)q * stride + i];
        }
        acc[0] += rho;
        for (int j = 0; j < 3; j++) acc[1 + j] += m[j];
        if (g) acc[4] += 0.5 * (e2 / cP.cs2 + gs);
        else acc[4] += 0.5 * (m[0] * m[0] + m[1] * m[1] + m[2] * m[2]) / rho;
    }
    for (int k = 0; k < 5; k++) sm[k][threadIdx.x] = acc[k];
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s)
            for (int k = 0; k < 5; k++) sm[k][threadIdx.x] += sm[k][threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x < 5) partial[blockIdx.x * 5 + threadIdx.x] = sm[threadIdx.x][0];
}

// ExponentialFilter<dim>::applyFilter (smoothing/ExponentialFilter.cpp:139-199) for the cells of ONE level, all Q populations
// of a distribution at once (CFDSolver::filter loops the populations, CFDSolver.cpp:866-869).  The reference visits the
// cells one after the other and a continuous FE shares face DoFs, so a cell reads what earlier cells wrote: the host sorts
// the cells into levels (a cell's level is above that of every earlier cell it shares a DoF with); the cells of one level
// share nothing and one launch handles them, one CTA per cell, thread i = row i of the two (p+1)^dim projections
// (FullMatrix::vmult: sum over j ascending).  toT / fromT are the transposed matrices, so a warp's weight loads coalesce;
// the cell's values sit in shared memory [n][Q] and are read as broadcasts.
template <int Q>
__global__ void __launch_bounds__(256)
k_filter_level(int n, const int32_t* __restrict__ cells, const int32_t* __restrict__ cell_dofs, const double* __restrict__ toT,
               const double* __restrict__ fromT, const double* __restrict__ sigma, double* __restrict__ pop, int64_t stride)
{
    extern __shared__ double smem_filter[];          // [n][Q]
    const int i = threadIdx.x;
    const int64_t cell = cells[blockIdx.x];
    int32_t idx = 0;
    if (i < n) {
        idx = __ldg(cell_dofs + cell * n + i);
#pragma unroll
        for (int q = 0; q < Q; q++) smem_filter[i * Q + q] = pop[(int64_t)q * stride + idx];
    }
    __syncthreads();
    double acc[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) acc[q] = 0.0;
    if (i < n) {
#pragma unroll 2
        for (int j = 0; j < n; j++) {
            const double w = __ldg(toT + (int64_t)j * n + i);
#pragma unroll
            for (int q = 0; q < Q; q++) acc[q] += w * smem_filter[j * Q + q];
        }
        const double sg = __ldg(sigma + i);
#pragma unroll
        for (int q = 0; q < Q; q++) acc[q] = sg * acc[q];
    }
    __syncthreads();
    if (i < n) {
#pragma unroll
        for (int q = 0; q < Q; q++) smem_filter[i * Q + q] = acc[q];
    }
    __syncthreads();
    if (i < n) {
#pragma unroll
        for (int q = 0; q < Q; q++) acc[q] = 0.0;
#pragma unroll 2
        for (int j = 0; j < n; j++) {
            const double w = __ldg(fromT + (int64_t)j * n + i);
#pragma unroll
            for (int q = 0; q < Q; q++) acc[q] += w * smem_filter[j * Q + q];
        }
#pragma unroll
        for (int q = 0; q < Q; q++) pop[(int64_t)q * stride + idx] = acc[q];
    }
}

static __global__ void k_conserved_final(int n_blocks, const double* __restrict__ partial, double* __restrict__ out)
{
    if (threadIdx.x < 5) {
        double s = 0.0;
        for (int b = 0; b < n_blocks; b++) s += partial[b * 5 + threadIdx.x];
        out[threadIdx.x] = s;
    }
}

