// tma_probe.cu -- stand-alone probe of the TMA path the grid kernels use (fp64 3-d tensor map, box load into shared memory
// completed on an mbarrier), with the tensor map (A) as a __grid_constant__ kernel parameter and (B) in global memory.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tma_probe tma_probe.cu   Run on the GPU box.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint64_t* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma3(double* dst, const void* tm, int x, int y, int z, uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(tm), "r"(x), "r"(y), "r"(z), "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

template <int MODE>
__global__ void probe(const __grid_constant__ CUtensorMap tm_param, const CUtensorMap* tm_global, int bx, int by, int bz, int vol, double* out)
{
    extern __shared__ __align__(128) double sm[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect(&bar, (unsigned)vol * 8);
        const void* tm = MODE == 0 ? (const void*)&tm_param : (const void*)tm_global;
        if (MODE == 2) asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tm) : "memory");
        tma3(sm, tm, bx, by, bz, &bar);
    }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < vol; i += blockDim.x) out[i] = sm[i];
}

typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv)
{
    const int NX = argc > 1 ? atoi(argv[1]) : 34, NY = argc > 2 ? atoi(argv[2]) : 33, NZ = argc > 3 ? atoi(argv[3]) : 5;
    const int BX = argc > 4 ? atoi(argv[4]) : 10, BY = argc > 5 ? atoi(argv[5]) : 5, BZ = argc > 6 ? atoi(argv[6]) : 4;
    const int OX = argc > 7 ? atoi(argv[7]) : 28, OY = argc > 8 ? atoi(argv[8]) : 30, OZ = argc > 9 ? atoi(argv[9]) : 2;
    printf("grid %dx%dx%d box %dx%dx%d at (%d,%d,%d)\n", NX, NY, NZ, BX, BY, BZ, OX, OY, OZ);
    std::vector<double> h((size_t)NX * NY * NZ);
    for (size_t i = 0; i < h.size(); i++) h[i] = (double)i;
    double *d, *out;
    cudaMalloc(&d, h.size() * 8);
    cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    cudaMalloc(&out, BX * BY * BZ * 8);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (!fn) { printf("no encode\n"); return 1; }
    CUtensorMap tm;
    const cuuint64_t gd[3] = {NX, NY, NZ}, gs[2] = {NX * 8, (cuuint64_t)NX * NY * 8};
    const cuuint32_t box[3] = {BX, BY, BZ}, es[3] = {1, 1, 1};
    CUresult r = ((Enc)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d\n", (int)r);
    CUtensorMap* dtm;
    cudaMalloc(&dtm, sizeof(tm));
    cudaMemcpy(dtm, &tm, sizeof(tm), cudaMemcpyHostToDevice);
    const int vol = BX * BY * BZ;
    for (int mode = 0; mode < 3; mode++) {
        cudaMemset(out, 0, vol * 8);
        if (mode == 0) probe<0><<<1, 128, vol * 8>>>(tm, dtm, OX, OY, OZ, vol, out);
        if (mode == 1) probe<1><<<1, 128, vol * 8>>>(tm, dtm, OX, OY, OZ, vol, out);
        if (mode == 2) probe<2><<<1, 128, vol * 8>>>(tm, dtm, OX, OY, OZ, vol, out);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<double> o((size_t)vol);
        cudaMemcpy(o.data(), out, vol * 8, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int z = 0; z < BZ; z++) for (int y = 0; y < BY; y++) for (int x = 0; x < BX; x++) {
            const int gx = OX + x, gy = OY + y, gz = OZ + z;
            const double want = (gx < NX && gy < NY && gz < NZ) ? (double)(((size_t)gz * NY + gy) * NX + gx) : 0.0;
            if (o[(size_t)((z * BY + y) * BX + x)] != want) bad++;
        }
        printf("mode %d (%s): %s, mismatches %d, o[0]=%g\n", mode, mode == 0 ? "grid_constant param" : mode == 1 ? "global" : "global+fence",
               cudaGetErrorString(e), bad, o[0]);
        if (e != cudaSuccess) break;
    }
    return 0;
}
