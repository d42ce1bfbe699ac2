// inst.cu -- one translation unit per stencil: compile with -DNB_D=<D> -DNB_Q=<Q> -DNB_NAME=<fn>
// [-DNB_WITH_G=1] [-DNB_FUSE_G=1].  Each unit owns its own copy of the constant block cP.
#include "launch.h"
#include "kernels.cuh"

#ifndef NB_WITH_G
#define NB_WITH_G 0
#endif
#ifndef NB_FUSE_G
#define NB_FUSE_G 0
#endif
#ifndef NB_FUSE_F
#define NB_FUSE_F 1
#endif

// entropic kernels exist where the reference has them: KBCStandard D2Q9 / D3Q15, MRTEntropic D3Q19
#if (NB_D == 2 && NB_Q == 9) || (NB_D == 3 && NB_Q == 15)
#define NB_IF_KBC(stmt) stmt
#else
#define NB_IF_KBC(stmt) return -1
#endif
#if (NB_D == 3 && NB_Q == 19)
#define NB_IF_MRTE(stmt) stmt
#else
#define NB_IF_MRTE(stmt) return -1
#endif
// collision_advanced rows of selectCollision: Regularized D2Q9 / D3Q15 / D3Q19, MultipleRelaxationTime D2Q9 / D3Q19
// (CollisionSelection.h:87-88,182-186)
#if (NB_D == 2 && NB_Q == 9) || (NB_D == 3 && NB_Q == 15) || (NB_D == 3 && NB_Q == 19)
#define NB_IF_REG(stmt) stmt
#else
#define NB_IF_REG(stmt) return -1
#endif
#if (NB_D == 2 && NB_Q == 9) || (NB_D == 3 && NB_Q == 19)
#define NB_IF_MRT(stmt) stmt
#define NB_HAS_MRT 1
#else
#define NB_IF_MRT(stmt) return -1
#define NB_HAS_MRT 0
#endif

namespace {

constexpr int D = NB_D;
constexpr int Q = NB_Q;

const void* s_owner = nullptr;
uint64_t s_version = 0;

int bind(const NbLaunch& L)
{
    if (s_owner != L.owner || s_version != L.version) {
        cudaError_t e = cudaMemcpyToSymbolAsync(cP, L.hc, sizeof(NbConst), 0, cudaMemcpyHostToDevice, L.stream);
        if (e != cudaSuccess) return (int)e;
        if (D == 3 && Q == 19 && L.mrt) {
            e = cudaMemcpyToSymbolAsync(cM, L.mrt, sizeof(NbMrtTables), 0, cudaMemcpyHostToDevice, L.stream);
            if (e != cudaSuccess) return (int)e;
        }
#if NB_HAS_MRT
        if (L.post_matrix) {
            e = cudaMemcpyToSymbolAsync(cA, L.post_matrix, sizeof(double) * NB_MRT_MAXQ * NB_MRT_MAXQ, 0, cudaMemcpyHostToDevice, L.stream);
            if (e != cudaSuccess) return (int)e;
        }
        if (L.mrt_std) {
            static_assert(sizeof(NbMrtStd) == sizeof(NbMrtStdHost), "MRT table layout");
            e = cudaMemcpyToSymbolAsync(cS, L.mrt_std, sizeof(NbMrtStd), 0, cudaMemcpyHostToDevice, L.stream);
            if (e != cudaSuccess) return (int)e;
        }
#endif
        if (L.grid_off && L.grid_off_dirs > 0) {
            e = cudaMemcpyToSymbolAsync(cGridOff, L.grid_off, sizeof(int16_t) * NB_GRID_MAXK * (size_t)L.grid_off_dirs, 0, cudaMemcpyHostToDevice, L.stream);
            if (e != cudaSuccess) return (int)e;
        }
        s_owner = L.owner;
        s_version = L.version;
    }
    return 0;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: set it whenever the current device has not seen it
template <class K>
int set_smem(K kernel, size_t bytes, unsigned* done_mask)
{
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 32 && ((*done_mask >> dev) & 1u)) return 0;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
    if (dev < 32) *done_mask |= 1u << dev;
    return 0;
}

inline unsigned grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }

int fused(const NbLaunch& L)
{
    int rc = bind(L);
    if (rc) return rc;
    const unsigned grid = L.grid_override ? L.grid_override : grid_for(L.A.n_slices * 32, 128);
    if (!L.with_g) {
#if NB_FUSE_F
#define NB_LAUNCH_F(EQ, FMT) k_stream_collide_f<D, Q, EQ, FMT><<<grid, 128, 0, L.stream>>>(L.A, L.xf, L.yf, L.rho, L.u, L.flag)
#define NB_LAUNCH_F_KIND(FMT)                                                                    \
    do {                                                                                         \
        if (L.eq == NB_EQ_BGK) NB_LAUNCH_F(NB_EQ_BGK, FMT);                                      \
        else if (L.eq == NB_EQ_QUARTIC) NB_LAUNCH_F(NB_EQ_QUARTIC, FMT);                         \
        else if (L.eq == NB_KIND_KBC) { NB_IF_KBC(NB_LAUNCH_F(NB_KIND_KBC, FMT)); }              \
        else if (L.eq == NB_KIND_MRT_ENTROPIC) { NB_IF_MRTE(NB_LAUNCH_F(NB_KIND_MRT_ENTROPIC, FMT)); } \
        else if (L.eq == NB_KIND_REGULARIZED) { NB_IF_REG(NB_LAUNCH_F(NB_KIND_REGULARIZED, FMT)); } \
        else if (L.eq == NB_KIND_MRT) { NB_IF_MRT(NB_LAUNCH_F(NB_KIND_MRT, FMT)); }              \
        else return -1;                                                                          \
    } while (0)
        if (L.fmt == NB_FMT_GRID) {
            // the attribute allows the largest buffers; the launch asks for what this context's tables were built for
            const size_t smem_gr = (size_t)(Q * NB_CTA_ROWS + 2 * NB_GRID_CAP) * sizeof(double);
            const size_t smem_use = (size_t)(Q * NB_CTA_ROWS + 2 * L.A.grid_cap) * sizeof(double);
            if (L.A.grid_cap < 16 || L.A.grid_cap > NB_GRID_CAP) return (int)cudaErrorInvalidValue;
#define NB_LAUNCH_FGR(EQ)                                                                                                   \
    do {                                                                                                                   \
        static unsigned attr_mask = 0;                                                                                     \
        int e = set_smem(k_stream_collide_f_grid<D, Q, EQ>, smem_gr, &attr_mask);                                          \
        if (e) return e;                                                                                                   \
        k_stream_collide_f_grid<D, Q, EQ><<<grid, NB_CTA_ROWS, smem_use, L.stream>>>(L.A, L.xf, L.yf, L.ygf, L.rho, L.u, L.flag); \
    } while (0)
            if (L.eq == NB_EQ_BGK) NB_LAUNCH_FGR(NB_EQ_BGK);
            else if (L.eq == NB_EQ_QUARTIC) NB_LAUNCH_FGR(NB_EQ_QUARTIC);
            else if (L.eq == NB_KIND_KBC) { NB_IF_KBC(NB_LAUNCH_FGR(NB_KIND_KBC)); }
            else if (L.eq == NB_KIND_MRT_ENTROPIC) { NB_IF_MRTE(NB_LAUNCH_FGR(NB_KIND_MRT_ENTROPIC)); }
            else if (L.eq == NB_KIND_REGULARIZED) { NB_IF_REG(NB_LAUNCH_FGR(NB_KIND_REGULARIZED)); }
            else if (L.eq == NB_KIND_MRT) { NB_IF_MRT(NB_LAUNCH_FGR(NB_KIND_MRT)); }
            else return -1;
#undef NB_LAUNCH_FGR
        }
        else if (L.fmt == NB_FMT_STAGED) {
            const size_t smem_st = (size_t)(Q * NB_CTA_ROWS + NB_STAGE_CAP) * sizeof(double);
#define NB_LAUNCH_FS(EQ)                                                                                                    \
    do {                                                                                                                   \
        static unsigned attr_mask = 0;                                                                                     \
        int e = set_smem(k_stream_collide_f_staged<D, Q, EQ>, smem_st, &attr_mask);                                        \
        if (e) return e;                                                                                                   \
        k_stream_collide_f_staged<D, Q, EQ><<<grid, NB_CTA_ROWS, smem_st, L.stream>>>(L.A, L.xf, L.yf, L.rho, L.u, L.flag); \
    } while (0)
            if (L.eq == NB_EQ_BGK) NB_LAUNCH_FS(NB_EQ_BGK);
            else if (L.eq == NB_EQ_QUARTIC) NB_LAUNCH_FS(NB_EQ_QUARTIC);
            else if (L.eq == NB_KIND_KBC) { NB_IF_KBC(NB_LAUNCH_FS(NB_KIND_KBC)); }
            else if (L.eq == NB_KIND_MRT_ENTROPIC) { NB_IF_MRTE(NB_LAUNCH_FS(NB_KIND_MRT_ENTROPIC)); }
            else if (L.eq == NB_KIND_REGULARIZED) { NB_IF_REG(NB_LAUNCH_FS(NB_KIND_REGULARIZED)); }
            else if (L.eq == NB_KIND_MRT) { NB_IF_MRT(NB_LAUNCH_FS(NB_KIND_MRT)); }
            else return -1;
#undef NB_LAUNCH_FS
        }
        else if (L.fmt == NB_FMT_DICT) NB_LAUNCH_F_KIND(NB_FMT_DICT);
        else NB_LAUNCH_F_KIND(NB_FMT_ELL);
#undef NB_LAUNCH_F_KIND
#undef NB_LAUNCH_F
#else
        return -1;
#endif
    } else {
#if NB_WITH_G && NB_FUSE_G
        const size_t smem_fg = (size_t)2 * Q * 128 * sizeof(double);
#define NB_LAUNCH_FG(EQ, FMT)                                                                                              \
    do {                                                                                                                   \
        static unsigned attr_mask = 0;                                                                                     \
        int e = set_smem(k_stream_collide_fg<D, Q, EQ, FMT>, smem_fg, &attr_mask);                                         \
        if (e) return e;                                                                                                   \
        k_stream_collide_fg<D, Q, EQ, FMT><<<grid, 128, smem_fg, L.stream>>>(L.A, L.xf, L.xg, L.yf, L.yg, L.rho, L.u, L.T, L.sensor, L.flag); \
    } while (0)
        if (L.fmt == NB_FMT_GRID) {
            const size_t smem_gr = (size_t)(2 * Q * NB_CTA_ROWS + 4 * NB_GRID_CAP_FGF) * sizeof(double);
#define NB_LAUNCH_FGGR(EQ)                                                                                                  \
    do {                                                                                                                   \
        static unsigned attr_mask = 0;                                                                                     \
        int e = set_smem(k_stream_collide_fg_grid<D, Q, EQ>, smem_gr, &attr_mask);                                         \
        if (e) return e;                                                                                                   \
        k_stream_collide_fg_grid<D, Q, EQ><<<grid, NB_CTA_ROWS, smem_gr, L.stream>>>(L.A, L.xf, L.xg, L.yf, L.yg, L.ygf, L.ygg, L.rho, L.u, L.T, L.sensor, L.flag); \
    } while (0)
            if (L.eq == NB_EQ_BGK) NB_LAUNCH_FGGR(NB_EQ_BGK); else NB_LAUNCH_FGGR(NB_EQ_QUARTIC);
#undef NB_LAUNCH_FGGR
        }
        else if (L.fmt == NB_FMT_STAGED) {
            const size_t smem_st = (size_t)(2 * Q * NB_CTA_ROWS + 2 * NB_STAGE_CAP_FG) * sizeof(double);
#define NB_LAUNCH_FGS(EQ)                                                                                                   \
    do {                                                                                                                   \
        static unsigned attr_mask = 0;                                                                                     \
        int e = set_smem(k_stream_collide_fg_staged<D, Q, EQ>, smem_st, &attr_mask);                                       \
        if (e) return e;                                                                                                   \
        k_stream_collide_fg_staged<D, Q, EQ><<<grid, NB_CTA_ROWS, smem_st, L.stream>>>(L.A, L.xf, L.xg, L.yf, L.yg, L.rho, L.u, L.T, L.sensor, L.flag); \
    } while (0)
            if (L.eq == NB_EQ_BGK) NB_LAUNCH_FGS(NB_EQ_BGK); else NB_LAUNCH_FGS(NB_EQ_QUARTIC);
#undef NB_LAUNCH_FGS
        }
        else if (L.fmt == NB_FMT_DICT) { if (L.eq == NB_EQ_BGK) NB_LAUNCH_FG(NB_EQ_BGK, NB_FMT_DICT); else NB_LAUNCH_FG(NB_EQ_QUARTIC, NB_FMT_DICT); }
        else { if (L.eq == NB_EQ_BGK) NB_LAUNCH_FG(NB_EQ_BGK, NB_FMT_ELL); else NB_LAUNCH_FG(NB_EQ_QUARTIC, NB_FMT_ELL); }
#undef NB_LAUNCH_FG
#else
        return -1;
#endif
    }
    return (int)cudaGetLastError();
}

int stream_grid(const NbLaunch& L)
{
    int rc = bind(L);
    if (rc) return rc;
    const unsigned grid = L.grid_override;
    if (grid == 0) return 0;
    if (L.n_rhs == 2) {
#if NB_WITH_G
        const size_t sm = (size_t)4 * NB_GRID_CAP_OF(Q, 2) * sizeof(double);
        static unsigned attr_mask = 0;
        int e = set_smem(k_stream_grid<D, Q, 2>, sm, &attr_mask);
        if (e) return e;
        k_stream_grid<D, Q, 2><<<grid, NB_CTA_ROWS, sm, L.stream>>>(L.A, L.xf, L.xg, L.yf, L.yg);
#else
        return -1;
#endif
    } else {
        const size_t sm = (size_t)2 * NB_GRID_CAP * sizeof(double);
        static unsigned attr_mask = 0;
        int e = set_smem(k_stream_grid<D, Q, 1>, sm, &attr_mask);
        if (e) return e;
        k_stream_grid<D, Q, 1><<<grid, NB_CTA_ROWS, sm, L.stream>>>(L.A, L.xf, nullptr, L.yf, nullptr);
    }
    return (int)cudaGetLastError();
}

int collide(const NbLaunch& L)
{
    int rc = bind(L);
    if (rc) return rc;
    const int64_t n = L.A.n_owned;
    const unsigned grid = grid_for(n, 128);
    if (!L.with_g) {
#define NB_LAUNCH_C(EQ)                                                                                               \
    do {                                                                                                              \
        if (L.force) k_collide_f<D, Q, EQ, true><<<grid, 128, 0, L.stream>>>(n, L.A.stride, L.yf, L.rho, L.u, L.in_init, L.flag);  \
        else k_collide_f<D, Q, EQ, false><<<grid, 128, 0, L.stream>>>(n, L.A.stride, L.yf, L.rho, L.u, L.in_init, L.flag);         \
    } while (0)
#define NB_LAUNCH_C_NOFORCE(EQ) k_collide_f<D, Q, EQ, false><<<grid, 128, 0, L.stream>>>(n, L.A.stride, L.yf, L.rho, L.u, L.in_init, L.flag)
        if (L.eq == NB_EQ_BGK) NB_LAUNCH_C(NB_EQ_BGK);
        else if (L.eq == NB_EQ_QUARTIC) NB_LAUNCH_C(NB_EQ_QUARTIC);
        else if (L.eq == NB_KIND_KBC) { NB_IF_KBC(NB_LAUNCH_C_NOFORCE(NB_KIND_KBC)); }
        else if (L.eq == NB_KIND_MRT_ENTROPIC) { NB_IF_MRTE(NB_LAUNCH_C_NOFORCE(NB_KIND_MRT_ENTROPIC)); }
        else if (L.eq == NB_KIND_REGULARIZED) { NB_IF_REG(NB_LAUNCH_C(NB_KIND_REGULARIZED)); }
        else if (L.eq == NB_KIND_MRT) { NB_IF_MRT(NB_LAUNCH_C(NB_KIND_MRT)); }
        else return -1;
#undef NB_LAUNCH_C
#undef NB_LAUNCH_C_NOFORCE
    } else {
#if NB_WITH_G
#define NB_LAUNCH_CG(EQ)                                                                                              \
    do {                                                                                                              \
        if (L.force) k_collide_fg<D, Q, EQ, true><<<grid, 128, 0, L.stream>>>(n, L.A.stride, L.yf, L.yg, L.rho, L.u, L.T, L.sensor, L.in_init, L.flag, L.gidx, L.ygf, L.ygg, L.A.gstride);  \
        else k_collide_fg<D, Q, EQ, false><<<grid, 128, 0, L.stream>>>(n, L.A.stride, L.yf, L.yg, L.rho, L.u, L.T, L.sensor, L.in_init, L.flag, L.gidx, L.ygf, L.ygg, L.A.gstride);         \
    } while (0)
        if (L.eq == NB_EQ_BGK) NB_LAUNCH_CG(NB_EQ_BGK);
        else NB_LAUNCH_CG(NB_EQ_QUARTIC);
#undef NB_LAUNCH_CG
#else
        return -1;
#endif
    }
    return (int)cudaGetLastError();
}

int conserved(const NbLaunch& L)
{
    int rc = bind(L);
    if (rc) return rc;
    k_conserved_partial<D, Q><<<L.n_partial_blocks, 256, 0, L.stream>>>(L.A.n_owned, L.A.stride, L.xf, L.with_g ? L.xg : nullptr, L.partial);
    k_conserved_final<<<1, 32, 0, L.stream>>>(L.n_partial_blocks, L.partial, L.out);
    return (int)cudaGetLastError();
}

int wall(const NbLaunch& L)
{
    int rc = bind(L);
    if (rc) return rc;
    if (L.n_hit_groups <= 0) return 0;
    k_wall_hits<D, Q><<<grid_for(L.n_hit_groups, 64), 64, 0, L.stream>>>(L.n_hit_groups, L.hit_group_dof, L.hit_group_off, L.hit_dir,
                                                                         L.hit_kind, L.hit_val, L.A.stride, L.yf, L.yg, L.gidx, L.ygg, L.A.gstride);
    return (int)cudaGetLastError();
}

#if NB_HAS_MRT
int post(const NbLaunch& L)
{
    int rc = bind(L);
    if (rc) return rc;
    const int64_t n = L.A.n_owned;
    k_post_matrix<Q><<<grid_for(n, 128), 128, 0, L.stream>>>(n, L.A.stride, L.yf);
    return (int)cudaGetLastError();
}
#endif

int filter(const NbLaunch& L)
{
    if (L.filt_n_cells <= 0) return 0;
    const int threads = (L.filt_n + 31) / 32 * 32;
    const size_t sm = (size_t)L.filt_n * Q * sizeof(double);
    // the size depends on the context's FE order: set per call (cheap), not once per device
    cudaError_t e = cudaFuncSetAttribute(k_filter_level<Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return (int)e;
    k_filter_level<Q><<<(unsigned)L.filt_n_cells, threads, sm, L.stream>>>(L.filt_n, L.filt_cells, L.filt_dofs, L.filt_toT, L.filt_fromT,
                                                                        L.filt_sigma, L.yf, L.A.stride);
    return (int)cudaGetLastError();
}

const NbStencilOps ops = {D, Q,
#if NB_FUSE_F || (NB_WITH_G && NB_FUSE_G)
                          fused,
#else
                          nullptr,
#endif
                          collide, conserved, wall, bind,
#if NB_HAS_MRT
                          post,
#else
                          nullptr,
#endif
                          stream_grid, filter};

}  // namespace

const NbStencilOps* NB_NAME() { return &ops; }
