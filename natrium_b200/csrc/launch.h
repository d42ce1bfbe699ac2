// launch.h -- host-side interface between the C-ABI unit (nb200.cu) and the per-stencil
// kernel units (inst.cu compiled once per (D,Q)).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "stream_common.cuh"

struct NbConst;
struct NbMrtHost;
struct NbMrtStdHost;

struct NbLaunch {
    cudaStream_t stream;
    StreamArgs A;
    const double* xf; const double* xg;   // current populations (fused: read)
    double* yf; double* yg;               // next populations (fused: write) / in-place buffers (collide)
    double* rho; double* u; double* T; double* sensor;
    int* flag;
    int eq;        // NB_EQ_BGK / NB_EQ_QUARTIC
    int fmt;       // NB_FMT_ELL / NB_FMT_DICT
    int with_g;
    int in_init;
    // constant-block ownership: the unit re-uploads when (owner, version) changed
    const NbConst* hc; const void* owner; uint64_t version;
    const NbMrtHost* mrt;   // MRTEntropic tables (D3Q19 unit only)
    const NbMrtStdHost* mrt_std;   // MultipleRelaxationTime tables (D2Q9 / D3Q19 units)
    const double* post_matrix;     // [19][19] host copy of the post-collision matrix (uploaded with the constants)
    int force;              // external-force hooks (stand-alone collide kernels)
    unsigned grid_override; // staged kernels over a CTA subset (A.cta_map): number of CTAs to launch, 0 = all
    // wall hits (k_wall_hits)
    int64_t n_hit_groups; const int32_t* hit_group_dof; const int64_t* hit_group_off;
    const int32_t* hit_dir; const int32_t* hit_kind; const double* hit_val;
    // conserved sums
    double* partial; int n_partial_blocks; double* out;
    // grid kernels (NB_FMT_GRID): grid copies the fused kernel writes next to yf / yg, and the host copy of the
    // per-direction offset table [(Q-1)][NB_GRID_MAXK] (uploaded into the unit's constant memory with the constants)
    double* ygf; double* ygg;
    const int32_t* gidx;    // collide with grid dual write: canonical row -> flat grid index (null = canonical only)
    const int16_t* grid_off; int grid_off_dirs;
    int n_rhs;              // stream_grid: 1 = the distribution in xf/yf, 2 = f and g in one pass
    // exponential filter (k_filter_level): one level of cells on the populations in yf
    int filt_n; int64_t filt_n_cells; const int32_t* filt_cells; const int32_t* filt_dofs;
    const double* filt_toT; const double* filt_fromT; const double* filt_sigma;
};

struct NbStencilOps {
    int D, Q;
    int (*fused)(const NbLaunch&);       // stream + collide in one kernel; nullptr if not built
    int (*collide)(const NbLaunch&);     // in-place collide
    int (*conserved)(const NbLaunch&);   // deterministic conserved sums
    int (*wall)(const NbLaunch&);        // wall hits on yf (and yg)
    int (*bind)(const NbLaunch&);        // uploads the constant block if this context's version is not the bound one
    int (*post)(const NbLaunch&);        // post-collision matrix on yf; nullptr where the reference has none
    int (*stream_grid)(const NbLaunch&); // stream only over the grid copy (xf[, xg] -> yf[, yg], canonical output)
    int (*filter)(const NbLaunch&);      // one level of the exponential filter on yf
};

const NbStencilOps* nb_ops_d2q9();
const NbStencilOps* nb_ops_d3q19();
const NbStencilOps* nb_ops_d3q15();
const NbStencilOps* nb_ops_d2q25();
const NbStencilOps* nb_ops_d3q45();
