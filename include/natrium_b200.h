/*
 * natrium_b200.h -- C ABI of libnatrium_b200: the B200-native implementation of NATriuM's
 * per-timestep hot path (semi-Lagrangian stream + collide, device-resident
 * DistributionFunctions, NCCL ghost exchange).
 *
 * Plain C, POD only: no C++/torch/deal.II types cross this boundary.  One context per MPI
 * rank / GPU; a context is NOT re-entrant (the reference is single-threaded per rank,
 * L/utilities/MPIGuard.cpp:14-16).  Every call returns NB200_OK (0) or a negative error
 * code; nb200_last_error() gives the text.  No exception crosses; the host shim re-throws
 * CollisionException / DensityZeroException on NB200_ERR_DENSITY.
 *
 * Citations are into /root/reference, L = src/library/natrium.
 *
 * Ownership: the host owns every pointer it passes in; the library copies before it
 * returns.  After an upload the device copy of the populations is authoritative until the
 * next download.  The streaming matrix is immutable between nb200_finalize_matrix() calls
 * (re-upload after SemiLagrangian::reassemble()/setDeltaT()).
 */
#ifndef NATRIUM_B200_H_
#define NATRIUM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nb200_ctx nb200_ctx;

enum nb200_status {
    NB200_OK = 0,
    NB200_ERR_ARG = -1,         /* bad argument / wrong call order */
    NB200_ERR_CUDA = -2,        /* CUDA runtime error (text in nb200_last_error) */
    NB200_ERR_NCCL = -3,        /* NCCL error */
    NB200_ERR_UNSUPPORTED = -4, /* "Collision model not implemented yet", CollisionSelection.h:102-110 */
    NB200_ERR_DENSITY = -5,     /* density < 1e-10 met in collide: CollisionException
                                   (AuxiliaryCollisionFunctions.h:53-56) / DensityZeroException
                                   (CollisionOperator.h:106-109) */
    NB200_ERR_NO_DEVICE = -6    /* no CUDA device: there is NO CPU fallback */
};

/* CollisionSchemeName / EquilibriumSchemeName subset on the hot path (L/utilities/ConfigNames.h) */
enum nb200_collision_scheme {
    NB200_BGK_STANDARD = 0,     /* collision_advanced BGKCollision, CollisionSchemes.h:17-119 */
    NB200_KBC_STANDARD = 1,     /* legacy KBCStandard::collideAll (D2Q9, D3Q15), L/collision/KBCStandard.cpp:88-1028 */
    NB200_MRT_ENTROPIC = 2,     /* legacy MRTEntropic::collideAllD3Q19, L/collision/MRTEntropic.cpp:167-305 */
    NB200_BGK_REGULARIZED = 3,  /* collision_advanced Regularized (D2Q9, D3Q15, D3Q19), CollisionSchemes.h:122-203 */
    NB200_MRT_STANDARD = 4      /* collision_advanced MultipleRelaxationTime (D2Q9, D3Q19), CollisionSchemes.h:206-265;
                                   needs nb200_set_mrt() */
};
/* ForceType (L/utilities/ConfigNames.h:114-119) */
enum nb200_force_type {
    NB200_NO_FORCING = 0,
    NB200_SHIFTING_VELOCITY = 1,
    NB200_EXACT_DIFFERENCE = 2,
    NB200_GUO = 3               /* "Force Type not implemented" in the reference (Aux...h:356-359) */
};
enum nb200_equilibrium_scheme {
    NB200_BGK_EQUILIBRIUM = 0,     /* Equilibria.h:17-83 */
    NB200_QUARTIC_EQUILIBRIUM = 1  /* Equilibria.h:87-267 */
};

/* Flat mirror of what GeneralCollisionData (AuxiliaryCollisionFunctions.h:105-202) pulls out
 * of SolverConfiguration / ProblemDescription / Stencil for the path. */
typedef struct nb200_collision_params {
    int32_t scheme;            /* nb200_collision_scheme */
    int32_t equilibrium;       /* nb200_equilibrium_scheme */
    int32_t with_g;            /* 0: selectCollision(f) overload; 1: selectCollision(f,g) overload */
    int32_t in_init;           /* inInitializationProcedure (CollisionOperator.h:79-91) */
    double viscosity;          /* problemDescription.getViscosity() */
    double dt;                 /* delta_t; tau = nu/(dt*cs2_scaled)+0.5 (Aux...h:63-68,177) */
    double gamma;              /* getHeatCapacityRatioGamma() (f+g only) */
    int32_t prandtl_set;       /* isPrandtlNumberSet() */
    int32_t sutherland_set;    /* isSutherlandLawSet() */
    double prandtl;            /* getPrandtlNumber() (default 1) */
    int32_t has_external_force; /* problemDescription.hasExternalForce() (CollisionOperator.h:83,93) */
    int32_t force_type;        /* nb200_force_type = configuration.getForcingScheme() */
    double force[3];           /* getExternalForce()->getForce() (GeneralCollisionData ctor, Aux...h:179-183) */
} nb200_collision_params;

/* ---- lifecycle ------------------------------------------------------------------------- */

/* Writes an opaque 128-byte NCCL unique id (rank 0 calls it, the host broadcasts it over its
 * own MPI/torch.distributed).  Replaces nothing in the reference (MPI_COMM_WORLD is implicit there). */
int nb200_get_unique_id(void *out128);

/* One context per rank; rank r <-> GPU `device`.  nccl_unique_id may be NULL when nranks==1.
 * Mirrors the per-rank objects built in CFDSolver<dim>::CFDSolver (L/solver/CFDSolver.cpp:77-395). */
int nb200_create(nb200_ctx **out, int device, int rank, int nranks, const void *nccl_unique_id);
void nb200_destroy(nb200_ctx *ctx);
const char *nb200_last_error(const nb200_ctx *ctx);

/* ---- static data: stencil, layout, streaming matrix, ghost plan -------------------------- */

/* Stencil tables: e is Q x D row-major *scaled* directions (Stencil::getDirections()), w the
 * weights, cs2_scaled = getSpeedOfSoundSquare() (L/stencils/Stencil.h:53-171). */
int nb200_set_stencil(nb200_ctx *ctx, int D, int Q, const double *e_scaled, const double *w,
                      double scaling, double cs2_scaled);

/* n_owned = |getLocallyOwnedDofs()|, n_ghost = |getLocallyRelevantDofs()| - n_owned
 * (L/advection/AdvectionOperator.h:106-112).  Local index space: [0,n_owned) owned, then ghosts.
 * with_g allocates the second distribution (CompressibleCFDSolver::m_g). */
int nb200_set_layout(nb200_ctx *ctx, int64_t n_owned, int64_t n_ghost, int with_g);

/* Optional locality hint, right after nb200_set_layout: the library stores owned DoF order[k] at internal
 * position k (populations, moments and matrix rows; ghosts stay behind the owned range).  Every other call keeps
 * speaking the caller's numbering -- uploads, downloads, CSR blocks and the halo plan are translated inside.
 * The natural order is the one SemiLagrangian::fillSparseObject itself walks: for each locally owned cell in
 * active-cell (p4est Morton) order, cell->get_dof_indices(), first visit wins
 * (L/advection/SemiLagrangian.cpp:201-223).  DoFs of one cell then sit next to each other, a warp reads one
 * source cell, and the gathered support values stay in L1. */
int nb200_set_dof_order(nb200_ctx *ctx, int64_t n_owned, const int32_t *order);

/* Optional structure hint, after nb200_set_layout (and nb200_set_dof_order, if used) and before the first
 * nb200_upload_block_csr: the local DoFs sit on a tensor-product grid.  coords[(n_owned + n_ghost)][dim] are the integer
 * grid coordinates of every local DoF (owned first, then the ghost slots), 0 <= coords[.][j] < dims[j]; two DoFs never
 * share a grid point.  A continuous FE_Q(p) space on any (stretched) hyper-rectangle mesh has such coordinates whatever its
 * DoF numbering: rank the distinct support-point coordinates of DoFTools::map_dofs_to_support_points per axis
 * (INTEGRATION.md).  fe_order = p tells the library where cells begin (grid coordinates 0, p, 2p, ... of the local grid are
 * cell faces; pass the coordinates shifted accordingly) or 0 if unknown.
 * With the hint the library keeps a second, lexicographic copy of the populations and the fused kernel fetches the
 * (p+1)^k support points of all rows of a tile as boxes of that grid with TMA tensor copies (cp.async.bulk.tensor);
 * rows whose columns do not form such a box (bounce-back blocks at walls, truncated rows) are taken from their dictionary
 * list.  Rows are summed in ascending grid position instead of the uploaded order.  Every other call is unchanged.
 * dim = 0 / coords = NULL removes the hint. */
int nb200_set_dof_grid(nb200_ctx *ctx, int dim, const int32_t *dims, const int32_t *coords, int fe_order);
/* out = { 1 if the grid kernels are in use, #tiles, rows taken from TMA boxes (summed over directions), rows taken from
 *         their dictionary list, #boxes (TMA copies per step and distribution), #passes, pass capacity (values), grid points } */
int nb200_grid_info(const nb200_ctx *ctx, int64_t out[8]);

/* One block (bi,bj) of getSystemMatrix() (distributed_sparse_block_matrix, (Q-1)x(Q-1) blocks,
 * SemiLagrangian.cpp:101,116-134) as local CSR: exactly what
 * block(bi,bj).trilinos_matrix().ExtractMyRowView gives row by row (idiom in
 * L/smoothing/VmultLimiter.cpp:32-60).  col < n_owned: owned column, else ghost slot. */
int nb200_upload_block_csr(nb200_ctx *ctx, int bi, int bj, int64_t n_rows, const int64_t *rowptr,
                           const int32_t *col_local, const double *val);

/* Builds the device streaming format from the uploaded blocks (after reassemble()). */
int nb200_finalize_matrix(nb200_ctx *ctx);

/* Device representation of getSystemMatrix().  Call before the first nb200_upload_block_csr.
 *   NB200_FORMAT_ELL   warp-sliced ELL, 12 B per stored entry, values bit-identical to the upload.
 *   NB200_FORMAT_DICT  (default) dictionary format: each row = (column-list id, weight-pattern id); rows that
 *                      read the same source cell share a list, rows whose values agree entry by entry to within
 *                      value_dedup_tol share a pattern.  value_dedup_tol = 0 keeps every value bit-identical;
 *                      the default 1e-14 is four orders below the 1e-10 below which the reference's own
 *                      assembly drops entries (L/advection/SemiLagrangian.cpp:483) and bounds the per-row
 *                      perturbation by K*1e-14*max|f| (K = row length <= (p+1)^dim).  With the grid hint
 *                      (nb200_set_dof_grid) and one distribution, two rows that are multiplied together (same position in two
 *                      neighbouring cells) and whose patterns differ by no more than twice the tolerance take the
 *                      first one's pattern, i.e. a stored weight is then within 3 * value_dedup_tol of its own.
 * The matrix the reference assembles on a regular mesh has only O((p+1)^dim) distinct rows per direction up
 * to round-off, which is what the dictionary exploits; on an unstructured mesh it degenerates to ELL.
 *   NB200_FORMAT_DICT_UNSTAGED  the same dictionary, but every row walks its own column list in global memory.
 *                      NB200_FORMAT_DICT drives the dictionary CTA by CTA instead: the distinct column lists of
 *                      128 consecutive rows are copied to shared memory once per pass and shared by the rows
 *                      ("staged"); it falls back to the unstaged kernels by itself when the rows of a CTA share
 *                      too little for the staged values of one direction to fit (see nb200_staging_info). */
enum nb200_matrix_format { NB200_FORMAT_ELL = 0, NB200_FORMAT_DICT = 1, NB200_FORMAT_DICT_UNSTAGED = 2 };
int nb200_set_matrix_format(nb200_ctx *ctx, int format, double value_dedup_tol);
/* out = { format, #weight patterns, #column lists, pool bytes, row-descriptor bytes, #row-length classes } */
int nb200_matrix_format_info(const nb200_ctx *ctx, int64_t out[6], double *value_dedup_tol);
/* out = { 1 if the staged kernels are in use, staged support values per step (all CTAs, all passes), #passes,
 *         largest pass (values), pass capacity (values) } */
int nb200_staging_info(const nb200_ctx *ctx, int64_t out[5]);

/* Ghost plan derived from the column map / IndexSets: for neighbour k, owned local indices
 * send_idx[send_off[k]..send_off[k+1]) go to rank nbr_rank[k]; ghost slots
 * [recv_off[k], recv_off[k+1]) (relative to n_owned) are filled from it.  Replaces the Epetra
 * Import inside vmult and DistributionFunctions::updateGhosted() (DistributionFunctions.h:282-293). */
int nb200_set_halo(nb200_ctx *ctx, int n_nbr, const int32_t *nbr_rank, const int64_t *send_off,
                   const int32_t *send_idx, const int64_t *recv_off);

/* Wall hits of the semi-Lagrangian boundary handler (SemiLagrangianBoundaryHandler::addHit,
 * L/boundaries/SemiLagrangianBoundaryHandler.cpp:14-23; BoundaryHit.h), flattened in HitList iteration order
 * (cell, point, hit -- the order SemiLagrangianBoundaryHandler::operate walks, :43-109).  Applied after the SpMV of
 * f inside nb200_stream(ctx, 0) and nb200_step, exactly where m_boundaryHandler.apply(f, f_old[, g], t) sits
 * (SemiLagrangian.h:163-169, CFDSolver.cpp:673, CompressibleCFDSolver.h:195).  The bounce itself is part of the
 * matrix (off-diagonal blocks); this is the per-hit correction:
 *   NB200_WALL_VELOCITY_NEQ_BOUNCE_BACK  value[h] = 2 w_dir rho (e_dir . u_wall(x_hit, t - dtHit)) / cs2 with rho = 1,
 *       evaluated by the host that owns the wall-velocity function (VelocityNeqBounceBack.cpp:137-195); it is added
 *       to f[dest_direction](dest_index).  Re-upload when the wall velocity changes in time.
 *   NB200_WALL_THERMAL_BOUNCE_BACK       value[h] = wall temperature (ThermalBounceBack.cpp:50-109; D3Q45 with g):
 *       f and g of the destination DoF are re-equilibrated to it (all 45 populations).
 * dest_index: owned local DoF (LagrangianPathDestination::index), dest_direction: its direction.  n_hits = 0 clears. */
enum nb200_wall_kind { NB200_WALL_VELOCITY_NEQ_BOUNCE_BACK = 0, NB200_WALL_THERMAL_BOUNCE_BACK = 1 };
int nb200_set_wall_hits(nb200_ctx *ctx, int64_t n_hits, const int32_t *dest_index, const int32_t *dest_direction,
                        const int32_t *kind, const double *value);

/* ---- DistributionFunctions storage (L/solver/DistributionFunctions.h:47-300) -------------- */

/* which: 0 = f, 1 = g.  q in [0,Q): f.at(q).  host: n_owned contiguous doubles, the array
 * ExtractView returns (CollisionOperator.h:38-48). */
int nb200_upload_population(nb200_ctx *ctx, int which, int q, const double *host, int64_t n);
int nb200_download_population(nb200_ctx *ctx, int which, int q, double *host, int64_t n);
/* All Q populations at once, host layout [Q][n_owned]. */
int nb200_upload_populations(nb200_ctx *ctx, int which, const double *host, int64_t n);
int nb200_download_populations(nb200_ctx *ctx, int which, double *host, int64_t n);
/* Same with page-locked host buffers and async copies on the context stream (e2e path). */
int nb200_upload_populations_async(nb200_ctx *ctx, int which, const double *pinned_host, int64_t n);
int nb200_download_populations_async(nb200_ctx *ctx, int which, double *pinned_host, int64_t n);

/* Initial macroscopic velocity for in_init collisions (u_raw is an input then). u: [D][n_owned]. */
int nb200_upload_velocity(nb200_ctx *ctx, const double *u, int64_t n);

/* Macroscopic density m_density (n_owned doubles).  Only MRTEntropic reads it: its guard looks at the density
 * stored by the previous call (L/collision/MRTEntropic.cpp:229-232).  A new layout starts with densities = 1. */
int nb200_upload_density(nb200_ctx *ctx, const double *rho, int64_t n);

/* ---- per-step operators ------------------------------------------------------------------ */

int nb200_set_collision(nb200_ctx *ctx, const nb200_collision_params *p);

/* Tables of MultipleRelaxationTime::SpecificCollisionData (CollisionSchemes.h:209-236): M = make_M(basis),
 * T = make_T(basis) (both Q x Q row-major), omega = make_diag(tau, basis, relax mode) (Q entries), all from
 * L/collision_advanced/AuxiliaryMRTFunctions.cpp.  Call before nb200_set_collision(scheme = NB200_MRT_STANDARD);
 * omega depends on tau, so call again when viscosity or dt change. */
int nb200_set_mrt(nb200_ctx *ctx, int Q, const double *M, const double *T, const double *omega);

/* DistributionFunctions::updateGhosted() for f (and g): NCCL neighbour exchange. No-op on 1 rank. */
int nb200_update_ghosted(nb200_ctx *ctx);

/* SemiLagrangian::stream / the vmult site: f.FStream = M * f_old.FStream, f0 untouched
 * (SemiLagrangian.h:150-161; CFDSolver.cpp:671-672; gStream CompressibleCFDSolver.h:279-314).
 * Includes the ghost refresh the Epetra column-map import performs inside vmult. */
int nb200_stream(nb200_ctx *ctx, int which);

/* selectCollision(...)->collideAll: in place on the current populations; writes rho, u (,T, sensor)
 * (CollisionOperator.h:26-224). */
int nb200_collide(nb200_ctx *ctx);

/* n_steps x { stream(f); [stream(g);] collide } fused into one pass over the populations per
 * step, device resident (CFDSolver::run loop body, CFDSolver.cpp:877-902 without output()). */
int nb200_step(nb200_ctx *ctx, int n_steps);

/* DataProcessor hook after collide (the loop over m_dataProcessors in CFDSolver::run, CFDSolver.cpp:889-891) for the one
 * processor on the path: PseudoEntropicStabilizer::apply (L/dataprocessors/PseudoEntropicStabilizer.cpp:152-290) replaces
 * the populations of every owned DoF by A f.  A (Q x Q, row-major) is the host's table (n/d, nd_d2q9_with_e or nd_d3q19,
 * :27-150); D2Q9 and D3Q19 only, as in the reference.  Once set it runs after the collision of every nb200_step
 * iteration; nb200_apply_post_collision applies it once to the current populations.  A = NULL removes it. */
int nb200_set_post_collision_matrix(nb200_ctx *ctx, int Q, const double *A);
int nb200_apply_post_collision(nb200_ctx *ctx);

/* Filter hook between stream and collide (CFDSolver::filter, L/solver/CFDSolver.cpp:859-874; compressibleFilter,
 * L/solver/CompressibleCFDSolver.h:315-333): ExponentialFilter<dim>::applyFilter (L/smoothing/ExponentialFilter.cpp:139-199)
 * on every population of f (and g).  The host keeps what the reference computes once with deal.II and hands it over:
 *   cell_dofs[n_cells][dofs_per_cell]  cell->get_dof_indices() of every locally owned cell IN THE ORDER of the reference's
 *                                      active-cell loop, as local indices (owned DoFs or ghost slots);
 *   to_legendre, from_legendre         getProjectToLegendre() / getProjectFromLegendre(), row-major [n][n], n = dofs_per_cell,
 *                                      in the same element-local numbering as cell_dofs;
 *   sigma[n]                           damping factor of every Legendre mode, exp(-alpha ((degree+1-Nc)/(max_degree+1-Nc))^s)
 *                                      for degree >= Nc and exactly 1 otherwise (the reference leaves those modes alone).
 * The reference's loop is sequential and cells of a continuous element share face DoFs: a cell reads what earlier cells
 * wrote.  The library keeps exactly that order -- cells are sorted into levels (above every earlier cell they share a DoF
 * with) and each level is one kernel launch -- so results equal the reference's whatever the cell order (lexicographic,
 * Morton, ...).  interval = getFilterInterval(): inside nb200_step the iteration counter m_i is advanced per step and the
 * filter runs when m_i % interval == 0 (those steps run stream, filter, collide as separate kernels); interval = 0 keeps
 * it out of nb200_step.  nb200_apply_filter applies it once to the current populations of distribution `which`
 * (the m_filter->applyFilter call site); nb200_set_iteration sets m_i (restart).  n_cells = 0 removes the filter.
 * Several ranks: a cell's ghost DoFs are read from the ghost slots and written there only locally (the reference writes a
 * non-ghosted vector); nb200_step refreshes all ghost slots of the streamed populations before the filter and the next
 * step's exchange overwrites them, as m_f.updateGhosted() does.  dofs_per_cell <= 256. */
int nb200_set_filter(nb200_ctx *ctx, int64_t n_cells, int dofs_per_cell, const int32_t *cell_dofs, const double *to_legendre,
                     const double *from_legendre, const double *sigma, int interval);
int nb200_apply_filter(nb200_ctx *ctx, int which);
int nb200_set_iteration(nb200_ctx *ctx, int64_t iteration);
/* out = { #cells, dofs per cell, #levels (kernel launches per distribution and application), interval, m_i } */
int nb200_filter_info(const nb200_ctx *ctx, int64_t out[5]);

/* One step driven with HOST buffers -- what a host-resident DistributionFunctions sees from
 * SemiLagrangian::stream(f_old, f, t) followed by selectCollision (SemiLagrangian.h:150-161, CollisionSelection.h:60-67):
 * f_in [Q][n] -> device, fused stream+collide, f_out [Q][n], rho [n], u [D][n] -> host (rho / u may be NULL).  Page-locked
 * buffers make the copies asynchronous; the call returns after enqueue, nb200_synchronize() fences it.  With
 * n_chunks > 1 upload, kernel and download are pipelined over pieces of the DoF range on separate streams (a piece's
 * kernel waits only for the pieces its rows read), so the two PCIe directions overlap; with several ranks the pieces that
 * hold values a neighbour needs are uploaded first, the ghost exchange runs behind them on its own stream and only the rows
 * that read ghost slots wait for it.  n_chunks <= 1, wall hits, f+g or a non-staged matrix format run the same legs in
 * sequence.  Results are identical either way. */
int nb200_step_host(nb200_ctx *ctx, const double *f_in, double *f_out, double *rho, double *u, int64_t n, int n_chunks);

/* The same for the compressible solver's two distributions (CompressibleCFDSolver::stream / gStream / collide,
 * L/solver/CompressibleCFDSolver.h:181-314): f_in, g_in [Q][n] -> device, one step, f_out, g_out [Q][n], rho [n], u [D][n],
 * T [n] -> host (rho / u / T may be NULL).  The legs run in sequence on the context stream. */
int nb200_step_host_fg(nb200_ctx *ctx, const double *f_in, const double *g_in, double *f_out, double *g_out, double *rho,
                       double *u, double *T, int64_t n);

/* ---- results ----------------------------------------------------------------------------- */

/* rho [n], u [D][n] (scaled, as written to m_velocity), T [n], sensor [n]; any pointer may be NULL. */
int nb200_download_moments(nb200_ctx *ctx, double *rho, double *u, double *T, double *sensor, int64_t n);

/* Global (all-rank) sums over owned DoFs of the current populations:
 * out = { sum rho, sum rho*u_x, sum rho*u_y, sum rho*u_z, sum energy }, momentum with the scaled velocities (sum_i e_i f_i).
 * energy, f only:  0.5 * |sum_i e_i f_i|^2 / rho  (kinetic energy, scaled velocities);
 * energy, f + g:   0.5 * (sum_i |e_i/scaling|^2 f_i / cs2 + sum_i g_i)  (the total energy the two distributions carry, in the
 *                  lattice units of calculateTemperature, AuxiliaryCollisionFunctions.h:290-307: = rho (0.5 |u|^2/cs2 + Cv T)).
 * The two branches are different quantities by design (the compressible one is what the f+g collision conserves); both are
 * plain sums over DoFs, a diagnostic mirror of PhysicalProperties::mass / kineticEnergy (L/solver/PhysicalProperties.cpp:29-131). */
int nb200_conserved(nb200_ctx *ctx, double out[5]);

/* Blocks until queued work is done; returns NB200_ERR_DENSITY if the sticky density flag is set. */
int nb200_synchronize(nb200_ctx *ctx);

/* ---- measurement helpers (CUDA events on the context's own stream) ------------------------ */
int nb200_timer_start(nb200_ctx *ctx);
int nb200_timer_stop(nb200_ctx *ctx, float *elapsed_ms);   /* synchronizes */
/* Number of kernels this library launched since context creation. */
int64_t nb200_kernel_launches(const nb200_ctx *ctx);
/* Writes a description of the device streaming format and its byte size. */
int nb200_matrix_info(const nb200_ctx *ctx, int64_t *nnz, int64_t *device_bytes, int64_t *padded_entries);
/* Raw cudaStream_t of the context (for profilers / external event timing). */
void *nb200_stream_handle(const nb200_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* NATRIUM_B200_H_ */
