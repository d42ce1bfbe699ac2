// kernels.cuh -- fused stream+collide and stand-alone collide kernels, templated on the stencil.
// Instantiated once per (D,Q) by inst.cu (one translation unit per stencil so that the
// build parallelises and each unit owns its constant-memory block).
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
#ifndef NB_STAGED_OCC_F
#define NB_STAGED_OCC_F 4
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
    const int64_t row = blockIdx.x * (int64_t)NB_CTA_ROWS + tid;
    const bool active = row < A.n_owned;
    tile[0][tid] = active ? x[row] : 0.0;
    const int p0 = __ldg(A.stage_cta + blockIdx.x), p1 = __ldg(A.stage_cta + blockIdx.x + 1);
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
#pragma unroll 1
        for (int a = ps.a0; a < ps.a1; a++) {
            const int2 d = reinterpret_cast<const int2*>(&tile[a + 1][tid])[0];
            double r, dummy;
            nb_row_dot_staged<1>(A, a, d, xs, xs, r, dummy);
            tile[a + 1][tid] = r;
        }
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
    const int64_t row = blockIdx.x * (int64_t)NB_CTA_ROWS + tid;
    const bool active = row < A.n_owned;
    tf[0][tid] = active ? xf[row] : 0.0;
    tg[0][tid] = active ? xg[row] : 0.0;
    const int p0 = __ldg(A.stage_cta + blockIdx.x), p1 = __ldg(A.stage_cta + blockIdx.x + 1);
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
#pragma unroll 1
        for (int a = ps.a0; a < ps.a1; a++) {
            const int2 d = reinterpret_cast<const int2*>(&tf[a + 1][tid])[0];
            double r0, r1;
            nb_row_dot_staged<2>(A, a, d, xsf, xsg, r0, r1);
            tf[a + 1][tid] = r0;
            tg[a + 1][tid] = r1;
        }
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

// Stand-alone collide (in place), f only.
template <int D, int Q, int EQ>
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
    if (nb_collide_f<D, Q, EQ>(f, rho, v, in_init != 0, EQ == NB_KIND_MRT_ENTROPIC ? rho_out[row] : 1.0)) *flag = 1;
#pragma unroll
    for (int q = 0; q < Q; q++) fbuf[(int64_t)q * stride + row] = f[q];
    rho_out[row] = rho;
    if (!in_init) {
#pragma unroll
        for (int j = 0; j < D; j++) u_out[(int64_t)j * n + row] = v[j];
    }
}

// Stand-alone collide (in place), f and g.
template <int D, int Q, int EQ>
__global__ void __launch_bounds__(128)
k_collide_fg(int64_t n, int64_t stride, double* __restrict__ fbuf, double* __restrict__ gbuf,
             double* __restrict__ rho_out, double* __restrict__ u_out, double* __restrict__ T_out,
             double* __restrict__ s_out, int in_init, int* __restrict__ flag)
{
    const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (row >= n) return;
    double f[Q], g[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) {
        f[q] = fbuf[(int64_t)q * stride + row];
        g[q] = gbuf[(int64_t)q * stride + row];
    }
    double rho, u[3], uo[3], T, sensor;
    if (in_init) {
#pragma unroll
        for (int j = 0; j < D; j++) uo[j] = u_out[(int64_t)j * n + row];
    }
    nb_collide_bgk_fg<D, Q, EQ>(f, g, rho, u, T, sensor, in_init ? uo : nullptr);
    if (rho < 1e-10) *flag = 1;
#pragma unroll
    for (int q = 0; q < Q; q++) {
        fbuf[(int64_t)q * stride + row] = f[q];
        gbuf[(int64_t)q * stride + row] = g[q];
    }
    rho_out[row] = rho;
    T_out[row] = T;
    s_out[row] = sensor;
    if (!in_init) {
#pragma unroll
        for (int j = 0; j < D; j++) u_out[(int64_t)j * n + row] = u[j] * cP.scaling;
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

static __global__ void k_conserved_final(int n_blocks, const double* __restrict__ partial, double* __restrict__ out)
{
    if (threadIdx.x < 5) {
        double s = 0.0;
        for (int b = 0; b < n_blocks; b++) s += partial[b * 5 + threadIdx.x];
        out[threadIdx.x] = s;
    }
}

