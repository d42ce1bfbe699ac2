// stream_common.cuh -- warp-sliced ELL row product shared by all streaming kernels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

struct StreamArgs {
    const double* __restrict__ ell_val;
    const int32_t* __restrict__ ell_idx;
    const int64_t* __restrict__ slice_off;   // [(Q-1)][n_slices+1]
    int64_t n_slices;
    int64_t n_owned;
    int64_t stride;
};

// one ELL row dot product for 1 or 2 right-hand sides (f and g share the matrix pass)
template <int NRHS>
__device__ __forceinline__ void nb_row_dot(const StreamArgs& A, int alpha_m1, int64_t slice, int lane,
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

