// B200Backend.h -- the C++ shim a NATriuM build adds next to L/solver/CFDSolver.h to run its stream + collide hot path on
// libnatrium_b200 (include/natrium_b200.h).  Header-only; it touches the reference's own types only through the members
// listed in shim/mock/Epetra_mock.h (Epetra_CrsMatrix / Epetra_BlockMap / Epetra_Import / Epetra_FEVector, deal.II's
// TrilinosWrappers::SparseMatrix::trilinos_matrix(), MPI::Vector::trilinos_vector(), natrium::DistributionFunctions::at(),
// Stencil getters), so the same file compiles against the real headers (define NATRIUM_B200_REAL_HEADERS and include
// CFDSolver.h first) and against the stand-ins that tests/cpp/shim_check.cpp uses in this repository.
//
// Reference sites replaced (L = src/library/natrium):
//   B200Backend ctor          getSystemMatrix() blocks -> device format      L/advection/SemiLagrangian.cpp:101,116-134
//   buildOwnedFirstNumbering  Epetra column map -> [owned | ghosts by owner]  (vmult's internal Import, CFDSolver.cpp:672)
//   haloPlanFromImporter      Epetra_Import export / remote lists -> nb200_set_halo
//   upload / download         DistributionFunctions::at(q) ExtractView idiom  L/collision_advanced/CollisionOperator.h:38-48
//   ensureHostMirror          lazy m_f mirror before output()/checkpoint      L/solver/CFDSolver.cpp:949-1101, Checkpoint.cpp:74-113
//   updateGhostedIfStale      DistributionFunctions::updateGhosted            L/solver/DistributionFunctions.h:282-293
//   stream / collide / step   CFDSolver::stream, collide, run loop            L/solver/CFDSolver.cpp:659-754,807-843,877-892
//   setWallHits               SemiLagrangianBoundaryHandler hit list          L/boundaries/SemiLagrangianBoundaryHandler.cpp:43-109
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "natrium_b200.h"

#ifndef NATRIUM_B200_REAL_HEADERS
#include "mock/Epetra_mock.h"
#endif

namespace natrium {

/// Local numbering the device library expects: owned DoFs [0, n_owned) in row-map order, then the ghost slots grouped by
/// owner rank (ascending), inside a rank in the order of the importer's remote list -- the order in which that rank's
/// exports arrive (Epetra sorts the remote IDs by owner before it builds the distributor, Epetra_Import.cpp).
struct OwnedFirstNumbering {
    int64_t n_owned = 0, n_ghost = 0;
    std::vector<int32_t> col2local;        // Epetra local column id -> device local index
    std::vector<int> ghost_gid, ghost_pid; // per ghost slot
};

inline OwnedFirstNumbering buildOwnedFirstNumbering(const Epetra_CrsMatrix& m)
{
    OwnedFirstNumbering N;
    const Epetra_Map& rows = m.RowMap();
    const Epetra_Map& cols = m.ColMap();
    N.n_owned = rows.NumMyElements();
    const int nc = cols.NumMyElements();
    N.col2local.assign((size_t)nc, -1);
    // remote columns in importer order (falls back to column-map order for a serial matrix without importer)
    std::vector<int> remote_cols;
    const Epetra_Import* imp = m.Importer();
    if (imp) remote_cols.assign(imp->RemoteLIDs(), imp->RemoteLIDs() + imp->NumRemoteIDs());
    std::vector<char> is_remote((size_t)nc, 0);
    for (int c : remote_cols) is_remote[(size_t)c] = 1;
    for (int c = 0; c < nc; c++) {
        if (is_remote[(size_t)c]) continue;
        const int lid = rows.LID(cols.GID(c));
        if (lid < 0) {                      // not in the importer's remote list and not owned: a serial matrix never has these
            remote_cols.push_back(c);
            is_remote[(size_t)c] = 1;
        } else {
            N.col2local[(size_t)c] = (int32_t)lid;
        }
    }
    const int nr = (int)remote_cols.size();
    std::vector<int> gids((size_t)nr), pids((size_t)nr, 0);
    for (int i = 0; i < nr; i++) gids[(size_t)i] = cols.GID(remote_cols[(size_t)i]);
    if (nr > 0) m.DomainMap().RemoteIDList(nr, gids.data(), pids.data(), nullptr);
    std::vector<int> order((size_t)nr);
    for (int i = 0; i < nr; i++) order[(size_t)i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return pids[(size_t)a] < pids[(size_t)b]; });
    N.n_ghost = nr;
    N.ghost_gid.resize((size_t)nr);
    N.ghost_pid.resize((size_t)nr);
    for (int s = 0; s < nr; s++) {
        const int i = order[(size_t)s];
        N.col2local[(size_t)remote_cols[(size_t)i]] = (int32_t)(N.n_owned + s);
        N.ghost_gid[(size_t)s] = gids[(size_t)i];
        N.ghost_pid[(size_t)s] = pids[(size_t)i];
    }
    return N;
}

/// Arguments of nb200_set_halo.
struct HaloPlan {
    std::vector<int32_t> nbr;
    std::vector<int64_t> send_off, recv_off;
    std::vector<int32_t> send_idx;
};

/// Neighbours = ranks we export to or import from; per neighbour the owned rows we send (the importer's export list of
/// that rank, in list order) and the number of ghost slots it fills (contiguous by construction of the numbering).
inline HaloPlan haloPlanFromImporter(const Epetra_Import* imp, const OwnedFirstNumbering& N)
{
    HaloPlan P;
    std::vector<int> ranks(N.ghost_pid.begin(), N.ghost_pid.end());
    const int ne = imp ? imp->NumExportIDs() : 0;
    for (int i = 0; i < ne; i++) ranks.push_back(imp->ExportPIDs()[i]);
    std::sort(ranks.begin(), ranks.end());
    ranks.erase(std::unique(ranks.begin(), ranks.end()), ranks.end());
    P.send_off.push_back(0);
    P.recv_off.push_back(0);
    for (int r : ranks) {
        P.nbr.push_back((int32_t)r);
        for (int i = 0; i < ne; i++)
            if (imp->ExportPIDs()[i] == r) P.send_idx.push_back((int32_t)imp->ExportLIDs()[i]);     // source-map LID = owned row
        P.send_off.push_back((int64_t)P.send_idx.size());
        int64_t cnt = 0;
        for (int p : N.ghost_pid) cnt += p == r;
        P.recv_off.push_back(P.recv_off.back() + cnt);
    }
    return P;
}

/// One CSR block in the device library's numbering (ExtractMyRowView idiom of L/smoothing/VmultLimiter.cpp:32-60).
struct LocalCsr {
    std::vector<int64_t> rowptr;
    std::vector<int32_t> col;
    std::vector<double> val;
};

inline LocalCsr extractBlock(const Epetra_CrsMatrix& B, const OwnedFirstNumbering& N, const Epetra_CrsMatrix& layout)
{
    // every block of the system matrix shares the row map; a block's own column map may be a subset of the layout block's
    LocalCsr out;
    const int n = B.NumMyRows();
    out.rowptr.assign((size_t)n + 1, 0);
    out.col.reserve((size_t)B.NumMyNonzeros());
    out.val.reserve((size_t)B.NumMyNonzeros());
    const bool same_cols = &B == &layout;
    for (int i = 0; i < n; i++) {
        double* v;
        int* idx;
        int k;
        B.ExtractMyRowView(i, k, v, idx);
        for (int j = 0; j < k; j++) {
            const int c = same_cols ? idx[j] : layout.ColMap().LID(B.ColMap().GID(idx[j]));
            if (c < 0) throw std::runtime_error("B200Backend: a block reads a column the layout block does not import");
            out.col.push_back(N.col2local[(size_t)c]);
            out.val.push_back(v[j]);
        }
        out.rowptr[(size_t)i + 1] = (int64_t)out.col.size();
    }
    return out;
}

#ifndef NATRIUM_B200_REAL_HEADERS
// L/utilities/ConfigNames.h, the enumerators the translation below names (same order as the reference)
enum CollisionSchemeName { BGK_STANDARD, BGK_STANDARD_TRANSFORMED, BGK_STEADY_STATE, BGK_MULTIPHASE, BGK_INCOMPRESSIBLE, MRT_STANDARD,
                           MRT_ENTROPIC, ENTROPIC_STABILIZED, KBC_STANDARD, KBC_CENTRAL, BGK_MULTI_AM4, BGK_MULTI_BDF2, BGK_REGULARIZED };
enum EquilibriumSchemeName { BGK_EQUILIBRIUM, QUARTIC_EQUILIBRIUM, INCOMPRESSIBLE_EQUILIBRIUM, STEADYSTATE_EQUILIBRIUM, ENTROPIC_EQUILIBRIUM };
enum ForceType { NO_FORCING, SHIFTING_VELOCITY, EXACT_DIFFERENCE, GUO };
#endif

/// CollisionSelection.h:60-122: the schemes on the path; everything else is "Collision model not implemented yet".
inline int toNb200(CollisionSchemeName s)
{
    switch (s) {
    case BGK_STANDARD: return NB200_BGK_STANDARD;
    case KBC_STANDARD: return NB200_KBC_STANDARD;
    case MRT_ENTROPIC: return NB200_MRT_ENTROPIC;
    case BGK_REGULARIZED: return NB200_BGK_REGULARIZED;
    case MRT_STANDARD: return NB200_MRT_STANDARD;
    default: throw CollisionException("Collision model not implemented yet (B200 backend)");
    }
}
inline int toNb200(EquilibriumSchemeName e)
{
    if (e == BGK_EQUILIBRIUM) return NB200_BGK_EQUILIBRIUM;
    if (e == QUARTIC_EQUILIBRIUM) return NB200_QUARTIC_EQUILIBRIUM;
    throw CollisionException("Equilibrium not implemented yet (B200 backend)");
}
inline int toNb200(ForceType f) { return (int)f; }      // same order: NO_FORCING, SHIFTING_VELOCITY, EXACT_DIFFERENCE, GUO

/// GeneralCollisionData's reads (AuxiliaryCollisionFunctions.h:105-202) from any configuration / problem description with
/// the reference's getter names.
template <class Config, class Problem>
nb200_collision_params collisionParams(Config& cfg, Problem& pd, double dt, bool with_g, bool in_init, size_t dim)
{
    nb200_collision_params p;
    std::memset(&p, 0, sizeof(p));
    p.scheme = toNb200(cfg.getCollisionScheme());
    p.equilibrium = toNb200(cfg.getEquilibriumScheme());
    p.with_g = with_g ? 1 : 0;
    p.in_init = in_init ? 1 : 0;
    p.viscosity = pd.getViscosity();
    p.dt = dt;
    p.gamma = cfg.getHeatCapacityRatioGamma();
    p.prandtl_set = cfg.isPrandtlNumberSet() ? 1 : 0;
    p.prandtl = cfg.getPrandtlNumber();
    p.sutherland_set = cfg.isSutherlandLawSet() ? 1 : 0;
    p.has_external_force = pd.hasExternalForce() ? 1 : 0;
    p.force_type = toNb200(cfg.getForcingScheme());
    if (pd.hasExternalForce())
        for (size_t i = 0; i < dim; i++) p.force[i] = pd.getExternalForce()->getForce()[i];
    return p;
}

class B200Backend {
    nb200_ctx* m_ctx = nullptr;
    OwnedFirstNumbering m_num;
    size_t m_Q = 0, m_dim = 0;
    bool m_withG = false;
    // which side holds the newest populations: the device after stream/collide/step, both after an upload or a download
    bool m_hostStale[2] = {false, false};
    bool m_ghostStale[2] = {false, false};
    int64_t m_downloads = 0;

    void check(int rc) const
    {
        if (rc == NB200_OK) return;
        const std::string msg = nb200_last_error(m_ctx);
        if (rc == NB200_ERR_DENSITY || rc == NB200_ERR_UNSUPPORTED) throw CollisionException(msg.c_str());
        throw std::runtime_error(msg);      // real build: natrium_errorexit (NATriuMException.h:61-76)
    }

public:
    /// After CFDSolver's constructor has run setupDoFs / setDeltaT / reassemble (CFDSolver.cpp:233-255).
    /// device: CUDA ordinal of this rank's GPU; nccl_unique_id: 128 bytes broadcast from rank 0 (nb200_get_unique_id), or
    /// null for a serial run.
    B200Backend(const distributed_sparse_block_matrix& M, const Stencil& st, bool with_g, int device, int rank, int nranks,
                const void* nccl_unique_id)
        : m_Q(st.getQ()), m_dim(st.getD()), m_withG(with_g)
    {
        int rc = nb200_create(&m_ctx, device, rank, nranks, nccl_unique_id);
        if (rc != NB200_OK) throw std::runtime_error("nb200_create failed: no CUDA device / NCCL (there is no CPU fallback)");
        try {
            std::vector<double> e(m_Q * m_dim), w(m_Q);
            for (size_t i = 0; i < m_Q; i++) {
                w[i] = st.getWeight(i);
                for (size_t d = 0; d < m_dim; d++) e[i * m_dim + d] = st.getDirection(i)(d);
            }
            check(nb200_set_stencil(m_ctx, (int)m_dim, (int)m_Q, e.data(), w.data(), st.getScaling(), st.getSpeedOfSoundSquare()));
            // layout from the block with the widest column map (the diagonal blocks all import the same ghost layer)
            const Epetra_CrsMatrix* layout = &M.block(0, 0).trilinos_matrix();
            for (size_t b = 1; b + 1 < m_Q; b++) {
                const Epetra_CrsMatrix& c = M.block(b, b).trilinos_matrix();
                if (c.ColMap().NumMyElements() > layout->ColMap().NumMyElements()) layout = &c;
            }
            m_num = buildOwnedFirstNumbering(*layout);
            check(nb200_set_layout(m_ctx, m_num.n_owned, m_num.n_ghost, with_g ? 1 : 0));
            for (size_t bi = 0; bi + 1 < m_Q; bi++)
                for (size_t bj = 0; bj + 1 < m_Q; bj++) {
                    const Epetra_CrsMatrix& B = M.block(bi, bj).trilinos_matrix();
                    if (B.NumMyNonzeros() == 0) continue;
                    const LocalCsr L = extractBlock(B, m_num, *layout);
                    check(nb200_upload_block_csr(m_ctx, (int)bi, (int)bj, m_num.n_owned, L.rowptr.data(), L.col.data(), L.val.data()));
                }
            check(nb200_finalize_matrix(m_ctx));
            if (nranks > 1) {
                const HaloPlan P = haloPlanFromImporter(layout->Importer(), m_num);
                check(nb200_set_halo(m_ctx, (int)P.nbr.size(), P.nbr.data(), P.send_off.data(), P.send_idx.data(), P.recv_off.data()));
            }
        } catch (...) {
            nb200_destroy(m_ctx);
            m_ctx = nullptr;
            throw;
        }
    }
    ~B200Backend() { nb200_destroy(m_ctx); }
    B200Backend(const B200Backend&) = delete;
    B200Backend& operator=(const B200Backend&) = delete;

    nb200_ctx* context() const { return m_ctx; }
    const OwnedFirstNumbering& numbering() const { return m_num; }
    int64_t downloads() const { return m_downloads; }

    /// Local index of an owned global DoF in a boundary hit (h.getDestination().index is global)
    template <class RowMap> int32_t localIndex(const RowMap& rows, int gid) const { return (int32_t)rows.LID(gid); }

    // ---- DistributionFunctions <-> device -------------------------------------------------------------------------
    void upload(DistributionFunctions& f, int which)
    {
        for (size_t q = 0; q < f.getQ(); q++) {
            double* p;
            int len;
            f.at(q).trilinos_vector().ExtractView(&p, &len);
            check(nb200_upload_population(m_ctx, which, (int)q, p, len));
        }
        m_hostStale[which] = false;
        m_ghostStale[which] = true;
    }
    void download(DistributionFunctions& f, int which)
    {
        for (size_t q = 0; q < f.getQ(); q++) {
            double* p;
            int len;
            f.at(q).trilinos_vector().ExtractView(&p, &len);
            check(nb200_download_population(m_ctx, which, (int)q, p, len));
        }
        f.updateGhosted();                  // host mirrors only; off the hot loop
        m_hostStale[which] = false;
        m_downloads++;
    }
    /// output() / Checkpoint::write / PhysicalProperties need the host vectors: copies only when the device has moved on
    /// since the last copy (CFDSolver.cpp:949-1101 calls this before it touches m_f, Checkpoint.cpp:74-113 likewise).
    void ensureHostMirror(DistributionFunctions& f, int which)
    {
        if (m_hostStale[which]) download(f, which);
    }
    /// DistributionFunctions::updateGhosted for the device copy: the steps exchange what they read by themselves, so this
    /// is only needed before something else reads ghost slots (the filter); a no-op while nothing changed.
    void updateGhostedIfStale()
    {
        if (!m_ghostStale[0] && !(m_withG && m_ghostStale[1])) return;
        check(nb200_update_ghosted(m_ctx));
        m_ghostStale[0] = m_ghostStale[1] = false;
    }
    /// Checkpoint::load: populations and the iteration counter the filter interval counts from
    void restart(DistributionFunctions& f, DistributionFunctions* g, int64_t iteration)
    {
        upload(f, 0);
        if (g) upload(*g, 1);
        check(nb200_set_iteration(m_ctx, iteration));
    }

    // ---- collision ---------------------------------------------------------------------------------------------------
    void setCollision(const nb200_collision_params& p) { check(nb200_set_collision(m_ctx, &p)); }
    void setMrt(const double* M, const double* T, const double* omega) { check(nb200_set_mrt(m_ctx, (int)m_Q, M, T, omega)); }
    void setStabilizer(const double* A) { check(nb200_set_post_collision_matrix(m_ctx, A ? (int)m_Q : 0, A)); }

    // ---- wall hits: the HitList flattened in the order SemiLagrangianBoundaryHandler::apply walks it ----------------
    void setWallHits(const std::vector<int32_t>& dest_index, const std::vector<int32_t>& dest_direction,
                     const std::vector<int32_t>& kind, const std::vector<double>& value)
    {
        check(nb200_set_wall_hits(m_ctx, (int64_t)dest_index.size(), dest_index.data(), dest_direction.data(), kind.data(), value.data()));
    }

    // ---- filter (ExponentialFilter tables from the host's own ExponentialFilter<dim> object) ------------------------
    void setFilter(const std::vector<int32_t>& cell_dofs, int dofs_per_cell, const double* to_legendre, const double* from_legendre,
                   const double* sigma, int interval)
    {
        check(nb200_set_filter(m_ctx, (int64_t)(cell_dofs.size() / (size_t)dofs_per_cell), dofs_per_cell, cell_dofs.data(), to_legendre,
                               from_legendre, sigma, interval));
    }

    // ---- the operators ------------------------------------------------------------------------------------------------
    void stream(int which)                  // CFDSolver::stream (which = 0) / CompressibleCFDSolver::gStream (which = 1)
    {
        check(nb200_stream(m_ctx, which));
        m_hostStale[which] = m_ghostStale[which] = true;
    }
    void collide()                          // CFDSolver::collide: throws CollisionException on density < 1e-10
    {
        check(nb200_collide(m_ctx));
        check(nb200_synchronize(m_ctx));
        m_hostStale[0] = m_ghostStale[0] = true;
        if (m_withG) m_hostStale[1] = m_ghostStale[1] = true;
    }
    void step(int n)                        // n iterations of the run() loop body, device resident
    {
        check(nb200_step(m_ctx, n));
        m_hostStale[0] = m_ghostStale[0] = true;
        if (m_withG) m_hostStale[1] = m_ghostStale[1] = true;
    }
    void sync() { check(nb200_synchronize(m_ctx)); }
    /// m_density / m_velocity / m_temperature mirrors (applyWriteableDensity / Velocity, CFDSolver.cpp:816-837)
    void moments(distributed_vector& rho, std::vector<distributed_vector>& u, distributed_vector* T = nullptr)
    {
        double *pr, *pT = nullptr;
        int len, l2;
        rho.trilinos_vector().ExtractView(&pr, &len);
        std::vector<double> ubuf((size_t)len * m_dim);
        if (T) T->trilinos_vector().ExtractView(&pT, &l2);
        check(nb200_download_moments(m_ctx, pr, ubuf.data(), pT, nullptr, len));
        for (size_t d = 0; d < m_dim; d++) {
            double* pu;
            u.at(d).trilinos_vector().ExtractView(&pu, &l2);
            std::memcpy(pu, ubuf.data() + d * (size_t)len, (size_t)len * sizeof(double));
        }
    }
};

}  // namespace natrium
