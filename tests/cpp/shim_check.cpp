// Drives shim/B200Backend.h -- the C++ shim a NATriuM build would add -- against stand-ins for the Epetra / deal.II classes
// (shim/mock/Epetra_mock.h).
//
//   shim_check halo <nx> <ny> <px> <py> <seed>
//       host logic only (no device call): a periodic nx x ny grid of DoFs with a 9-point stencil, partitioned into px x py
//       blocks (up to 8 neighbours, edge and corner ghosts -- what a p4est Z-curve partition gives); every rank gets a mock
//       Epetra_CrsMatrix whose column map and importer are laid out the way Epetra lays them out (owned columns first, remote
//       columns in arbitrary order, remote list sorted by owner with arbitrary order inside an owner, exports answering the
//       receivers' lists).  Checks buildOwnedFirstNumbering / haloPlanFromImporter: owned-first numbering, ghosts contiguous
//       per owner, and for every ordered pair (A, B) the DoFs A sends to B are exactly B's ghost slots for A, in order.
//   shim_check run <dump>
//       the real library through the shim on one GPU: matrix blocks, stencil, initial populations and the oracle's
//       populations after `steps` steps come from a dump written by tests/test_gpu_parity.py; the mock matrix gets a
//       permuted column map so that col2local is exercised; prints the max relative error.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <random>
#include <string>

#include "../../shim/B200Backend.h"

using namespace natrium;

static int halo_mode(int nx, int ny, int px, int py, unsigned seed)
{
    std::mt19937 rng(seed);
    const int R = px * py, N = nx * ny;
    std::vector<int> owner((size_t)N);
    for (int y = 0; y < ny; y++) for (int x = 0; x < nx; x++) owner[(size_t)(y * nx + x)] = (y * py / ny) * px + (x * px / nx);
    struct Rank { Epetra_CrsMatrix m; Epetra_Import imp; std::vector<int> remote_gids_sorted; OwnedFirstNumbering num; HaloPlan plan; };
    std::vector<Rank> ranks((size_t)R);
    for (int r = 0; r < R; r++) {
        Rank& K = ranks[(size_t)r];
        Epetra_CrsMatrix& m = K.m;
        for (int g = 0; g < N; g++) if (owner[(size_t)g] == r) m.row_map.gids.push_back(g);
        m.row_map.finish();
        m.row_map.owner_of_gid = &owner;
        // columns: owned first (row-map order), remote ones shuffled
        std::vector<int> remote;
        std::vector<char> seen((size_t)N, 0);
        for (int g : m.row_map.gids) {
            const int x = g % nx, y = g / nx;
            for (int dy = -1; dy <= 1; dy++) for (int dx = -1; dx <= 1; dx++) {
                const int c = ((y + dy + ny) % ny) * nx + (x + dx + nx) % nx;
                if (owner[(size_t)c] != r && !seen[(size_t)c]) { seen[(size_t)c] = 1; remote.push_back(c); }
            }
        }
        std::shuffle(remote.begin(), remote.end(), rng);
        m.col_map.gids = m.row_map.gids;
        m.col_map.gids.insert(m.col_map.gids.end(), remote.begin(), remote.end());
        m.col_map.finish();
        m.rowptr.push_back(0);
        for (int g : m.row_map.gids) {
            const int x = g % nx, y = g / nx;
            for (int dy = -1; dy <= 1; dy++) for (int dx = -1; dx <= 1; dx++) {
                const int c = ((y + dy + ny) % ny) * nx + (x + dx + nx) % nx;
                m.indices.push_back(m.col_map.LID(c));
                m.values.push_back(1.0 / 9.0);
            }
            m.rowptr.push_back((int)m.indices.size());
        }
        // importer: remote list grouped by owner, arbitrary order inside an owner (Epetra's sort is not stable)
        std::vector<int> rl;
        for (int c = m.row_map.NumMyElements(); c < m.col_map.NumMyElements(); c++) rl.push_back(c);
        std::shuffle(rl.begin(), rl.end(), rng);
        std::stable_sort(rl.begin(), rl.end(), [&](int a, int b) { return owner[(size_t)m.col_map.GID(a)] < owner[(size_t)m.col_map.GID(b)]; });
        K.imp.n_same = m.row_map.NumMyElements();
        K.imp.remote_lids = rl;
        for (int c : rl) K.remote_gids_sorted.push_back(m.col_map.GID(c));
        m.importer = R > 1 ? &K.imp : nullptr;
    }
    // exports: what every receiver asks of us, in the receiver's order, receivers ascending
    for (int a = 0; a < R; a++)
        for (int b = 0; b < R; b++) {
            if (a == b) continue;
            for (int g : ranks[(size_t)b].remote_gids_sorted)
                if (owner[(size_t)g] == a) { ranks[(size_t)a].imp.export_lids.push_back(ranks[(size_t)a].m.row_map.LID(g)); ranks[(size_t)a].imp.export_pids.push_back(b); }
        }
    int max_nbr = 0;
    for (int r = 0; r < R; r++) {
        Rank& K = ranks[(size_t)r];
        K.num = buildOwnedFirstNumbering(K.m);
        K.plan = haloPlanFromImporter(K.m.Importer(), K.num);
        const OwnedFirstNumbering& Nn = K.num;
        if (Nn.n_owned != K.m.NumMyRows() || Nn.n_ghost != K.m.ColMap().NumMyElements() - K.m.NumMyRows()) { printf("FAIL sizes on rank %d\n", r); return 1; }
        std::vector<char> used((size_t)(Nn.n_owned + Nn.n_ghost), 0);
        for (int c = 0; c < K.m.ColMap().NumMyElements(); c++) {
            const int32_t l = Nn.col2local[(size_t)c];
            if (l < 0 || l >= Nn.n_owned + Nn.n_ghost || used[(size_t)l]) { printf("FAIL col2local is not a bijection on rank %d\n", r); return 1; }
            used[(size_t)l] = 1;
            const int gid = K.m.ColMap().GID(c);
            if (l < Nn.n_owned ? K.m.RowMap().GID(l) != gid : Nn.ghost_gid[(size_t)(l - Nn.n_owned)] != gid) { printf("FAIL col2local gid mismatch on rank %d\n", r); return 1; }
        }
        for (size_t s = 1; s < Nn.ghost_pid.size(); s++) if (Nn.ghost_pid[s] < Nn.ghost_pid[s - 1]) { printf("FAIL ghosts not grouped by owner\n"); return 1; }
        // CSR in device numbering replays the 9-point average of the gid field
        const LocalCsr L = extractBlock(K.m, Nn, K.m);
        for (int64_t i = 0; i < Nn.n_owned; i++) {
            double acc = 0.0, ref = 0.0;
            for (int64_t k = L.rowptr[(size_t)i]; k < L.rowptr[(size_t)i + 1]; k++) {
                const int32_t l = L.col[(size_t)k];
                acc += L.val[(size_t)k] * (l < Nn.n_owned ? K.m.RowMap().GID(l) : Nn.ghost_gid[(size_t)(l - Nn.n_owned)]);
            }
            const int g = K.m.RowMap().GID((int)i), x = g % nx, y = g / nx;
            for (int dy = -1; dy <= 1; dy++) for (int dx = -1; dx <= 1; dx++) ref += (((y + dy + ny) % ny) * nx + (x + dx + nx) % nx) / 9.0;
            if (std::fabs(acc - ref) > 1e-9) { printf("FAIL extractBlock row %lld on rank %d\n", (long long)i, r); return 1; }
        }
        max_nbr = std::max(max_nbr, (int)K.plan.nbr.size());
    }
    // pairwise: A's sends to B == B's ghost slots for A, in order
    int64_t exchanged = 0;
    for (int a = 0; a < R; a++) {
        const Rank& A = ranks[(size_t)a];
        for (size_t ia = 0; ia < A.plan.nbr.size(); ia++) {
            const int b = A.plan.nbr[ia];
            const Rank& B = ranks[(size_t)b];
            size_t ib = 0;
            while (ib < B.plan.nbr.size() && B.plan.nbr[ib] != a) ib++;
            if (ib == B.plan.nbr.size()) { printf("FAIL neighbour lists not symmetric (%d, %d)\n", a, b); return 1; }
            const int64_t ns = A.plan.send_off[ia + 1] - A.plan.send_off[ia], nr = B.plan.recv_off[ib + 1] - B.plan.recv_off[ib];
            if (ns != nr) { printf("FAIL %d sends %lld to %d which expects %lld\n", a, (long long)ns, b, (long long)nr); return 1; }
            for (int64_t k = 0; k < ns; k++) {
                const int gid_sent = A.m.RowMap().GID(A.plan.send_idx[(size_t)(A.plan.send_off[ia] + k)]);
                const int gid_slot = B.num.ghost_gid[(size_t)(B.plan.recv_off[ib] + k)];
                if (gid_sent != gid_slot || B.num.ghost_pid[(size_t)(B.plan.recv_off[ib] + k)] != a) { printf("FAIL order of the exchange %d -> %d at %lld\n", a, b, (long long)k); return 1; }
            }
            exchanged += ns;
        }
    }
    printf("OK ranks=%d max_neighbours=%d exchanged=%lld\n", R, max_nbr, (long long)exchanged);
    return 0;
}

struct MockConfig {
    CollisionSchemeName scheme = BGK_STANDARD;
    CollisionSchemeName getCollisionScheme() { return scheme; }
    EquilibriumSchemeName getEquilibriumScheme() { return BGK_EQUILIBRIUM; }
    double getHeatCapacityRatioGamma() { return 1.4; }
    bool isPrandtlNumberSet() { return false; }
    double getPrandtlNumber() { return 1.0; }
    bool isSutherlandLawSet() { return false; }
    ForceType getForcingScheme() { return NO_FORCING; }
};
struct MockForce { const double* getForce() const { static const double z[3] = {0, 0, 0}; return z; } };
struct MockProblem {
    double nu;
    double getViscosity() { return nu; }
    bool hasExternalForce() { return false; }
    const MockForce* getExternalForce() { static MockForce f; return &f; }
};

static int run_mode(const char* path)
{
    FILE* fp = fopen(path, "rb");
    if (!fp) { printf("FAIL cannot open %s\n", path); return 1; }
    int64_t h[5];
    double par[4];
    if (fread(h, 8, 5, fp) != 5 || fread(par, 8, 4, fp) != 4) return 1;
    const int D = (int)h[0], Q = (int)h[1], steps = (int)h[4];
    const int64_t n = h[2];
    Stencil st;
    st.D = (size_t)D; st.Q = (size_t)Q; st.scaling = par[0]; st.cs2 = par[1];
    st.e.resize((size_t)Q * D); st.w.resize((size_t)Q);
    if (fread(st.e.data(), 8, st.e.size(), fp) != st.e.size() || fread(st.w.data(), 8, st.w.size(), fp) != st.w.size()) return 1;
    // mock system matrix: diagonal blocks, column map = a permutation of the row map (serial: no importer)
    std::mt19937 rng(5);
    distributed_sparse_block_matrix M;
    M.nb = (size_t)(Q - 1);
    M.blocks.resize(M.nb * M.nb);
    std::vector<int> colperm((size_t)n);
    std::iota(colperm.begin(), colperm.end(), 0);
    for (size_t b = 0; b < M.nb * M.nb; b++) {
        Epetra_CrsMatrix& m = M.blocks[b].m;
        m.row_map.gids.resize((size_t)n);
        std::iota(m.row_map.gids.begin(), m.row_map.gids.end(), 100);       // global ids need not start at 0
        m.row_map.finish();
        m.col_map = m.row_map;
        m.rowptr.assign((size_t)n + 1, 0);
    }
    for (int a = 0; a < Q - 1; a++) {
        int64_t nnz;
        if (fread(&nnz, 8, 1, fp) != 1) return 1;
        std::vector<int64_t> rp((size_t)n + 1);
        std::vector<int32_t> col((size_t)nnz);
        std::vector<double> val((size_t)nnz);
        if (fread(rp.data(), 8, rp.size(), fp) != rp.size() || fread(col.data(), 4, col.size(), fp) != col.size() || fread(val.data(), 8, val.size(), fp) != val.size()) return 1;
        Epetra_CrsMatrix& m = M.block((size_t)a, (size_t)a).m;
        if (a % 2 == 1) {                   // every other block: its own, permuted column map
            std::shuffle(colperm.begin(), colperm.end(), rng);
            for (int64_t c = 0; c < n; c++) m.col_map.gids[(size_t)colperm[(size_t)c]] = 100 + (int)c;
            m.col_map.finish();
        }
        m.rowptr.assign(rp.begin(), rp.end());
        m.indices.resize((size_t)nnz);
        for (int64_t k = 0; k < nnz; k++) m.indices[(size_t)k] = m.col_map.LID(100 + col[(size_t)k]);
        m.values = val;
    }
    DistributionFunctions f;
    f.m_f.resize((size_t)Q);
    std::vector<double> expect((size_t)Q * n);
    for (int q = 0; q < Q; q++) { f.m_f[(size_t)q].ev.v.resize((size_t)n); if (fread(f.m_f[(size_t)q].ev.v.data(), 8, (size_t)n, fp) != (size_t)n) return 1; }
    if (fread(expect.data(), 8, expect.size(), fp) != expect.size()) return 1;
    fclose(fp);

    B200Backend backend(M, st, false, 0, 0, 1, nullptr);
    MockConfig cfg;
    MockProblem pd{par[2]};
    backend.setCollision(collisionParams(cfg, pd, par[3], false, false, (size_t)D));
    backend.upload(f, 0);
    backend.ensureHostMirror(f, 0);                       // nothing moved yet: no copy
    if (backend.downloads() != 0) { printf("FAIL mirror copied although it was current\n"); return 1; }
    // reference order for the first step, fused for the rest
    backend.stream(0);
    backend.collide();
    backend.step(steps - 1);
    backend.sync();
    backend.ensureHostMirror(f, 0);
    backend.ensureHostMirror(f, 0);                       // second call is free
    if (backend.downloads() != 1 || f.ghost_updates != 1) { printf("FAIL lazy mirror: %lld downloads\n", (long long)backend.downloads()); return 1; }
    double err = 0.0;
    for (int q = 0; q < Q; q++)
        for (int64_t i = 0; i < n; i++) {
            const double e = expect[(size_t)(q * n + i)];
            err = std::max(err, std::fabs(f.m_f[(size_t)q].ev.v[(size_t)i] - e) / std::fabs(e));
        }
    distributed_vector rho;
    rho.ev.v.resize((size_t)n);
    std::vector<distributed_vector> u((size_t)D);
    for (auto& v : u) v.ev.v.resize((size_t)n);
    backend.moments(rho, u);
    double mass = 0.0;
    for (double r : rho.ev.v) mass += r;
    // an unsupported scheme surfaces as the reference's exception type
    bool threw = false;
    try { cfg.scheme = BGK_MULTIPHASE; backend.setCollision(collisionParams(cfg, pd, par[3], false, false, (size_t)D)); } catch (CollisionException&) { threw = true; }
    if (!threw) { printf("FAIL unsupported scheme did not throw CollisionException\n"); return 1; }
    printf("%s max_rel_err=%.3e mean_density=%.12f steps=%d n=%lld\n", err <= 1e-12 ? "OK" : "FAIL", err, mass / (double)n, steps, (long long)n);
    return err <= 1e-12 ? 0 : 1;
}

int main(int argc, char** argv)
{
    if (argc >= 7 && std::string(argv[1]) == "halo") return halo_mode(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), (unsigned)atoi(argv[6]));
    if (argc >= 3 && std::string(argv[1]) == "run") return run_mode(argv[2]);
    printf("usage: shim_check halo nx ny px py seed | shim_check run dump\n");
    return 2;
}
