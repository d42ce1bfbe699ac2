"""GPU parity tests: the CUDA path, called through the C ABI (ctypes -> libnatrium_b200.so), against
the CPU oracle on identical inputs.

Bar (BASELINE.json north_star): <= 1e-12 relative per step in double precision for f (and g);
conserved sums (mass, momentum, energy) agree to <= 1e-13 relative over 1000 steps.
"""
import numpy as np
import pytest

from tests import common
from tests.common import rel_err

pytestmark = pytest.mark.gpu

TOL_STEP = 1e-12      # north_star: 1e-12 relative per step
TOL_CONS = 1e-13      # conserved sums vs oracle


# device formats of the streaming matrix: library default (dictionary, value tolerance 1e-14), bit-faithful
# dictionary (tolerance 0) and the generic warp-sliced ELL
# (format, value tolerance, cell-blocked internal DoF order)
# a 4th element "grid": the host also gives the structure hint nb200_set_dof_grid -> TMA box kernels (NB_FMT_GRID)
FORMATS = [None, ("dict", 1e-14, True), ("dict", 0.0, False), ("ell", 0.0, False), ("ell", 0.0, True),
           ("dict-unstaged", 1e-14, False), ("dict-unstaged", 1e-14, True), ("dict", 1e-14, False, "grid"), ("dict", 0.0, True, "grid")]
FORMAT_IDS = ["dict-default", "dict-cellorder", "dict-exact", "ell", "ell-cellorder", "dict-unstaged", "dict-unstaged-cellorder",
              "grid", "grid-cellorder-exact"]


def fake_grid_hint(ctx, n, dim=2):
    """A grid hint for DoFs that sit on no mesh at all (random CSR inputs): any injective placement is valid input, the
    rows then match no box template and the grid kernels take every row from its dictionary list."""
    nx = int(np.ceil(n ** (1.0 / dim)))
    i = np.arange(n)
    if dim == 2:
        coords, dims = np.stack([i % nx, i // nx], axis=1), [nx, (n + nx - 1) // nx]
    else:
        coords, dims = np.stack([i % nx, (i // nx) % nx, i // (nx * nx)], axis=1), [nx, nx, (n + nx * nx - 1) // (nx * nx)]
    ctx.set_dof_grid(dims, coords, 0)


def _fmt_code(name):
    from natrium_b200 import _capi
    return {"dict": _capi.FORMAT_DICT, "ell": _capi.FORMAT_ELL, "dict-unstaged": _capi.FORMAT_DICT_UNSTAGED}[name]


def make_ctx(case, with_matrix=True, fmt=None):
    from natrium_b200 import Context, harness, _capi
    c, st, pb, dt = common.product_problem(case)
    ctx = Context(0)
    ctx.set_stencil(st.getDirections(), st.getWeights(), st.getScaling(), st.getSpeedOfSoundSquare())
    part = harness.SlabPartition(pb, st, dt)
    ctx.set_layout(part.n_owned, part.n_ghost, bool(c.get("with_g")))
    if fmt is not None:
        ctx.set_matrix_format(_fmt_code(fmt[0]), fmt[1])
        if fmt[2]:
            ctx.set_dof_order(part.cell_blocked_order())
        if len(fmt) > 3:
            ctx.set_dof_grid(*part.grid_coords(), fe_order=pb.p)
    if with_matrix:
        harness.upload_streaming_matrix(ctx, pb, part, st, dt)
    return ctx, c, st, pb, dt, part


def set_collision(ctx, c, dt, **kw):
    from natrium_b200 import _capi
    if c.get("with_g"):
        ctx.set_collision(c["nu"], dt, equilibrium=_capi.QUARTIC_EQUILIBRIUM, with_g=True, gamma=1.4,
                          prandtl=kw.get("prandtl", 0.71), sutherland=kw.get("sutherland", True))
    else:
        ctx.set_collision(c["nu"], dt, equilibrium=kw.get("equilibrium", _capi.BGK_EQUILIBRIUM))


@pytest.mark.parametrize("fmt", FORMATS, ids=FORMAT_IDS)
@pytest.mark.parametrize("case", ["c1_tgv2d_d2q9", "tgv3d_d3q19_small", "tgv3d_d3q15", "tgv2d_d2q25", "tgv3d_d3q45"])
def test_stream_matches_oracle(case, fmt, oracle_lib):
    """nb200_stream == vmult (CFDSolver.cpp:671-672): f.FStream = M f_old.FStream, f0 untouched."""
    o = common.oracle_problem(case)
    ctx, c, st, pb, dt, part = make_ctx(case, fmt=fmt)
    assert dt == o["dt"]
    ctx.upload_populations(0, o["f"])
    ctx.stream(0)
    got = ctx.download_populations(0)
    ref = oracle_lib.stream(o["blocks"], o["f"])
    assert np.array_equal(got[0], o["f"][0])
    assert rel_err(got, ref) <= TOL_STEP
    info = ctx.matrix_info()
    assert info["nnz"] == sum(b.nnz for b in o["blocks"].values())
    if c.get("with_g"):
        ctx.upload_populations(1, o["g"])
        ctx.stream(1)
        assert rel_err(ctx.download_populations(1), oracle_lib.stream(o["blocks"], o["g"])) <= TOL_STEP
    ctx.close()


def test_staged_kernels_selected():
    """The default dictionary format drives the staged kernels on the meshes of the path and falls back to the
    per-row kernels when told to (FORMAT_DICT_UNSTAGED) or when rows share too little (random sparse block)."""
    import scipy.sparse as sp
    from natrium_b200 import Context, Stencil
    for case in ["c1_tgv2d_d2q9", "tgv3d_d3q19_small", "tgv3d_d3q45"]:
        for fmt, want in [(None, True), (("dict", 1e-14, True), True), (("dict-unstaged", 1e-14, True), False)]:
            ctx = make_ctx(case, fmt=fmt)[0]
            info = ctx.matrix_format_info()
            assert info["staged"] == want, (case, fmt, info)
            if want:
                assert 0 < info["largest_pass"] <= info["pass_capacity"]
            ctx.close()
    st = Stencil("D2Q9", 1.0)
    n = 1000
    ctx = Context(0)
    ctx.set_stencil(st.getDirections(), st.getWeights(), 1.0, st.getSpeedOfSoundSquare())
    ctx.set_layout(n, 0, False)
    m = sp.random(n, n, density=0.15, random_state=np.random.default_rng(1), format="csr", dtype=np.float64)
    ctx.upload_block_csr(0, 0, m.indptr, m.indices, m.data)
    ctx.finalize_matrix()
    assert not ctx.matrix_format_info()["staged"]
    ctx.close()


def test_constant_streaming():
    """SemiLagrangian2D/3D_ConstantStreaming_test (SemiLagrangian_test.cpp:519-596): M*1 = 1."""
    for case in ["c1_tgv2d_d2q9", "tgv3d_d3q19_p2"]:
        ctx, c, st, pb, dt, part = make_ctx(case)
        ones = np.ones((st.getQ(), part.n_owned))
        ctx.upload_populations(0, ones)
        ctx.stream(0)
        got = ctx.download_populations(0)
        assert np.sum((got - ones) ** 2) <= 1e-6          # the reference's own bound
        assert np.max(np.abs(got - ones)) <= 1e-13        # and what fp64 actually gives
        ctx.close()


def synthetic_populations(Q, n):
    """f_i(j) of BGKStandard_test.cpp:373-380 (integer division makes i/(i+1) = 0)."""
    i = np.arange(Q, dtype=np.float64)[:, None]
    j = np.arange(n, dtype=np.float64)[None, :]
    return 1.5 + np.sin(1.5 * i) + 0.001 + (0.5 * np.cos(j)) ** 2 + 0 * i


@pytest.mark.parametrize("stencil,eq", [("D2Q9", 0), ("D2Q9", 1), ("D3Q19", 0), ("D3Q15", 0), ("D2Q25H", 1), ("D3Q45", 0), ("D3Q45", 1)])
def test_collide_f_matches_oracle(stencil, eq, oracle_lib):
    """selectCollision(f) rows of CollisionSelection.h on the BGKStandard_test population; tau=0.9+0.5-ish, dt=0.1."""
    from natrium_b200 import Context, Stencil
    scaling = 1.0 if stencil in ("D2Q25H", "D3Q45") else 2.5
    st = Stencil(stencil, scaling)
    ost = oracle_lib.Stencil(stencil, scaling)
    n, dt = 1000, 0.1
    nu = 0.9 * dt * st.getSpeedOfSoundSquare()
    f = synthetic_populations(st.getQ(), n) * st.getWeights()[:, None]
    ctx = Context(0)
    ctx.set_stencil(st.getDirections(), st.getWeights(), st.getScaling(), st.getSpeedOfSoundSquare())
    ctx.set_layout(n, 0, False)
    ctx.set_collision(nu, dt, equilibrium=eq)
    ctx.upload_populations(0, f)
    ctx.collide()
    ctx.synchronize()
    got = ctx.download_populations(0)
    rho, u = ctx.download_moments()
    ref = f.copy()
    # reference quirk: f-only D3Q45 + QUARTIC_EQUILIBRIUM dispatches to BGKEquilibrium (CollisionSelection.h:199)
    rrho, ru, rc = oracle_lib.collide_bgk(ost, ref, nu, dt, equilibrium=0 if stencil == "D3Q45" else eq)
    assert rc == 0
    assert rel_err(got, ref) <= TOL_STEP
    assert rel_err(rho, rrho) <= 1e-14
    assert np.max(np.abs(u - ru)) <= 1e-13 * max(1.0, np.max(np.abs(ru)))
    # collision invariants (BGKStandardCollisionInvariants_test): mass and momentum conserved
    assert np.max(np.abs(got.sum(axis=0) - f.sum(axis=0))) <= 1e-13 * np.max(f.sum(axis=0))
    ctx.close()


@pytest.mark.parametrize("stencil,eq,prandtl,sutherland", [("D2Q25H", 1, 0.71, True), ("D2Q25H", 1, None, False),
                                                           ("D2Q25H", 0, None, False), ("D3Q45", 1, 0.71, True),
                                                           ("D3Q45", 1, None, False)])
def test_collide_fg_matches_oracle(stencil, eq, prandtl, sutherland, oracle_lib):
    """selectCollision(f, g): relaxWithG with quartic equilibrium, Prandtl correction, Sutherland law."""
    from natrium_b200 import Context, Stencil, harness
    st = Stencil(stencil, 1.0)
    ost = oracle_lib.Stencil(stencil, 1.0)
    n, dt, nu, gamma = 777, 0.05, 0.002, 1.4
    rng = np.random.default_rng(7)
    rho = 1.0 + 0.1 * rng.standard_normal(n)
    u = 0.1 * rng.standard_normal((st.getD(), n))
    T = 1.0 + 0.05 * rng.standard_normal(n)
    f, g = harness.quartic_equilibrium_distributions(st, rho, u, T, gamma)
    f *= 1.0 + 0.01 * rng.standard_normal(f.shape)      # push off equilibrium
    g *= 1.0 + 0.01 * rng.standard_normal(g.shape)
    ctx = Context(0)
    ctx.set_stencil(st.getDirections(), st.getWeights(), st.getScaling(), st.getSpeedOfSoundSquare())
    ctx.set_layout(n, 0, True)
    if sutherland:
        ctx.set_dof_order(rng.permutation(n))          # internal DoF order must be invisible to the caller
    ctx.set_collision(nu, dt, equilibrium=eq, with_g=True, gamma=gamma, prandtl=prandtl, sutherland=sutherland)
    ctx.upload_populations(0, f)
    ctx.upload_populations(1, g)
    ctx.collide()
    ctx.synchronize()
    gf, gg = ctx.download_populations(0), ctx.download_populations(1)
    rho_d, u_d, T_d, s_d = ctx.download_moments(want_T=True)
    rf, rg = f.copy(), g.copy()
    rrho, ru, rT, rs, rc = oracle_lib.collide_bgk_fg(ost, rf, rg, nu, dt, equilibrium=eq, gamma=gamma, prandtl=prandtl,
                                                     sutherland=sutherland)
    assert rc == 0
    assert rel_err(gf, rf) <= TOL_STEP
    assert rel_err(gg, rg) <= TOL_STEP
    assert rel_err(T_d, rT) <= 1e-12 and rel_err(rho_d, rrho) <= 1e-14
    assert np.max(np.abs(s_d - rs)) <= 1e-12 * np.max(np.abs(rs))
    ctx.close()


@pytest.mark.parametrize("fmt", FORMATS, ids=FORMAT_IDS)
@pytest.mark.parametrize("case", ["c1_tgv2d_d2q9", "tgv3d_d3q19_small", "tgv3d_d3q15", "tgv2d_d2q25", "tgv3d_d3q45"])
def test_fused_step_matches_oracle_per_step(case, fmt, oracle_lib):
    """10 steps of nb200_step(1) vs the reference-ordered CPU step from the same state each step."""
    o = common.oracle_problem(case)
    ctx, c, st, pb, dt, part = make_ctx(case, fmt=fmt)
    with_g = bool(c.get("with_g"))
    set_collision(ctx, c, dt)
    stepper = oracle_lib.ReferenceOrderStepper(o["st"], o["blocks"], o["dofs"].N, c["nu"], dt,
                                               equilibrium=1 if with_g else 0, with_g=with_g, gamma=1.4,
                                               prandtl=0.71 if with_g else None, sutherland=with_g)
    f = o["f"].copy()
    g = o["g"].copy() if with_g else None
    ctx.upload_populations(0, f)
    if with_g:
        ctx.upload_populations(1, g)
    worst = 0.0
    for s in range(10):
        ctx.step(1)
        ctx.synchronize()
        assert stepper.step(f, g) == 0
        got = ctx.download_populations(0)
        e = rel_err(got, f)
        if with_g:
            e = max(e, rel_err(ctx.download_populations(1), g))
        worst = max(worst, e)
        assert e <= TOL_STEP, (case, s, e)
        # identical state for the next step (per-step comparison)
        ctx.upload_populations(0, f)
        if with_g:
            ctx.upload_populations(1, g)
    rho, u = ctx.download_moments()[:2]
    assert rel_err(rho, stepper.rho) <= 1e-13
    print(case, "worst per-step rel err", worst)
    ctx.close()


def test_unfused_equals_fused():
    """nb200_stream + nb200_collide (reference order) == nb200_step (fused)."""
    o = common.oracle_problem("tgv3d_d3q19_p2")
    ctx, c, st, pb, dt, part = make_ctx("tgv3d_d3q19_p2")
    set_collision(ctx, c, dt)
    ctx.upload_populations(0, o["f"])
    ctx.step(3)
    a = ctx.download_populations(0)
    ctx.upload_populations(0, o["f"])
    for _ in range(3):
        ctx.stream(0)
        ctx.collide()
    b = ctx.download_populations(0)
    assert rel_err(a, b) <= 1e-15
    ctx.close()


@pytest.mark.parametrize("fmt", FORMATS, ids=FORMAT_IDS)
def test_conserved_moments_1000_steps(fmt, oracle_lib):
    """Config 1 (TGV2D, D2Q9, p=4, 8x8 cells): sums of rho, rho*u, energy vs the oracle at steps 1, 10, 100, 1000,
    and the physics bound of integration test #11 (E_kin(t)/E_kin(0) = exp(-4 nu t))."""
    case = "c1_tgv2d_d2q9"
    o = common.oracle_problem(case)
    ctx, c, st, pb, dt, part = make_ctx(case, fmt=fmt)
    set_collision(ctx, c, dt)
    stepper = oracle_lib.ReferenceOrderStepper(o["st"], o["blocks"], o["dofs"].N, c["nu"], dt)
    f = o["f"].copy()
    oracle_lib.collide_bgk(o["st"], f, c["nu"], dt)          # run(): collide once before the loop
    ctx.upload_populations(0, o["f"])
    ctx.collide()
    e = o["st"].e

    def sums(fa):
        rho = fa.sum(axis=0)
        m = e.T @ fa
        return np.array([rho.sum(), m[0].sum(), m[1].sum(), 0.0, (0.5 * (m ** 2).sum(axis=0) / rho).sum()])

    E0 = sums(f)[4]
    done = 0
    for target in (1, 10, 100, 1000):
        ctx.step(target - done)
        for _ in range(target - done):
            stepper.step(f)
        done = target
        got, ref = ctx.conserved(), sums(f)
        assert abs(got[0] - ref[0]) <= TOL_CONS * abs(ref[0]), (target, got, ref)
        assert abs(got[4] - ref[4]) <= 1e-11 * abs(ref[4]), (target, got, ref)   # energy ~1e-3 of mass scale
        scale = np.abs(e).max() * ref[0]                     # momentum sums cancel to ~0: absolute vs mass scale
        assert np.max(np.abs(got[1:3] - ref[1:3])) <= TOL_CONS * scale, (target, got, ref)
        assert rel_err(ctx.download_populations(0), f) <= 1e-10, target   # trajectories stay together
    ratio = ctx.conserved()[4] / E0
    assert abs(ratio - np.exp(-4 * c["nu"] * 1000 * dt)) < 1e-2
    ctx.synchronize()
    ctx.close()


def test_uniform_flow_stays_uniform():
    """CFDSolver_SteadyStreaming_test (CFDSolver_test.cpp:44-112): |rho-1|, |u-0.1| < 1e-5 after 100 steps."""
    from natrium_b200 import harness
    ctx, c, st, pb, dt, part = make_ctx("tgv3d_d3q19_p2")
    set_collision(ctx, c, dt)
    n = part.n_owned
    rho, u = np.ones(n), np.full((3, n), 0.1)
    ctx.upload_populations(0, harness.equilibrium_distributions(st, rho, u))
    ctx.step(100)
    ctx.synchronize()
    r, v = ctx.download_moments()
    assert np.max(np.abs(r - 1)) < 1e-5 and np.max(np.abs(v - 0.1)) < 1e-5
    assert np.max(np.abs(r - 1)) < 1e-12          # it is in fact a fixed point to round-off
    ctx.close()


@pytest.mark.parametrize("ordered", [False, True])
def test_in_initialization_collide(ordered, oracle_lib):
    """inInitializationProcedure: velocities are an input (CollisionOperator.h:79-91)."""
    from natrium_b200 import Context, Stencil
    st, ost = Stencil("D2Q9", 1.0), oracle_lib.Stencil("D2Q9", 1.0)
    n = 300
    f = synthetic_populations(9, n) * st.getWeights()[:, None]
    u0 = 0.05 * np.vstack([np.sin(np.arange(n)), np.cos(np.arange(n))])
    ctx = Context(0)
    ctx.set_stencil(st.getDirections(), st.getWeights(), 1.0, st.getSpeedOfSoundSquare())
    ctx.set_layout(n, 0, False)
    if ordered:
        ctx.set_dof_order(np.random.default_rng(1).permutation(n))
        with pytest.raises(Exception):
            ctx.set_dof_order(np.zeros(n, dtype=np.int32))     # not a permutation
    ctx.set_collision(0.03, 0.1, in_init=True)
    ctx.upload_populations(0, f)
    ctx.upload_velocity(u0)
    ctx.collide()
    got = ctx.download_populations(0)
    ref = f.copy()
    _, ru, _ = oracle_lib.collide_bgk(ost, ref, 0.03, 0.1, in_init=True, u_init=u0.copy())
    assert rel_err(got, ref) <= TOL_STEP
    assert np.array_equal(ctx.download_moments()[1], u0)       # velocities untouched
    ctx.close()


def test_error_behaviour():
    """CollisionException on rho < 1e-10; 'not implemented' for unsupported models; argument errors."""
    from natrium_b200 import CollisionException, Context, NatriumB200Error, Stencil, _capi
    st = Stencil("D2Q9", 1.0)
    ctx = Context(0)
    with pytest.raises(NatriumB200Error):
        ctx.set_layout(10, 0)                                   # stencil not set yet
    ctx.set_stencil(st.getDirections(), st.getWeights(), 1.0, st.getSpeedOfSoundSquare())
    ctx.set_layout(64, 0, False)
    with pytest.raises(CollisionException):
        ctx.set_collision(0.1, 0.1, scheme=_capi.MRT_ENTROPIC)  # MRTEntropic has no D2Q9 implementation
    ctx.set_collision(0.1, 0.1)
    ctx.upload_populations(0, np.zeros((9, 64)))
    ctx.collide()
    with pytest.raises(CollisionException) as ei:
        ctx.synchronize()
    assert ei.value.code == _capi.NB200_ERR_DENSITY
    with pytest.raises(NatriumB200Error):
        ctx.step(1)                                             # no matrix
    bad = st.getDirections()[::-1].copy()
    with pytest.raises(CollisionException):
        ctx.set_stencil(bad, st.getWeights(), 1.0, st.getSpeedOfSoundSquare())   # hard-coded D2Q9 order violated
    ctx.close()


@pytest.mark.parametrize("fmt", FORMATS, ids=FORMAT_IDS)
def test_ragged_and_offdiagonal_blocks(fmt, oracle_lib):
    """Generic CSR input: ragged rows, empty rows, off-diagonal (wall-bounce) blocks, duplicates-free;
    n not a multiple of the slice size.  Block (5,5) has more distinct row lengths than the dictionary
    format has exact classes (padded power-of-two classes take over)."""
    import scipy.sparse as sp
    from natrium_b200 import Context, Stencil, _capi
    st = Stencil("D2Q9", 1.0)
    n = 1000 + 7
    rng = np.random.default_rng(3)
    blocks = {}
    for (bi, bj, dens) in [(0, 0, 0.01), (0, 2, 0.002), (1, 1, 0.02), (3, 1, 0.001), (4, 4, 0.0), (7, 7, 0.05), (7, 0, 0.01),
                           (5, 5, 0.15)]:
        m = sp.random(n, n, density=dens, random_state=rng, format="csr", dtype=np.float64)
        m.sort_indices()
        blocks[(bi, bj)] = m
    assert len(set(np.diff(blocks[(5, 5)].indptr).tolist())) > 50
    ctx = Context(0)
    ctx.set_stencil(st.getDirections(), st.getWeights(), 1.0, st.getSpeedOfSoundSquare())
    ctx.set_layout(n, 0, False)
    if fmt is not None:
        ctx.set_matrix_format(_fmt_code(fmt[0]), fmt[1])
        if fmt[2]:
            ctx.set_dof_order(rng.permutation(n))
        if len(fmt) > 3:
            fake_grid_hint(ctx, n)
    for (bi, bj), m in blocks.items():
        ctx.upload_block_csr(bi, bj, m.indptr, m.indices, m.data)
    ctx.finalize_matrix()
    f = rng.standard_normal((9, n))
    ctx.upload_populations(0, f)
    ctx.stream(0)
    got = ctx.download_populations(0)
    ref = oracle_lib.stream(blocks, f)
    for q in range(1, 9):
        if not any(k[0] == q - 1 for k in blocks):
            ref[q] = 0.0          # vmult of an empty block row gives 0
    assert np.max(np.abs(got - ref)) <= 1e-13 * np.max(np.abs(ref))
    ctx.close()


def test_host_mirror_solver(oracle_lib):
    """The reference-shaped host API (CFDSolver / selectCollision / SemiLagrangian.stream) drives the same path."""
    from natrium_b200 import CFDSolver, SolverConfiguration, harness
    case = "tgv2d_small"
    o = common.oracle_problem(case)
    c = o["c"]
    cfg = SolverConfiguration()
    cfg.setStencil("D2Q9")
    cfg.setStencilScaling(c["scaling"])
    cfg.setSedgOrderOfFiniteElement(c["p"])
    cfg.setCFL(c["cfl"])
    pb = harness.CartesianProblem(c["dim"], c["cells"], c["p"])
    solver = CFDSolver(cfg, pb, c["nu"])
    x = solver.getAdvectionOperator().getPartition().owned_points()
    assert np.array_equal(x, o["x"])
    solver.setInitialFields(o["rho"], o["u"])
    stepper = oracle_lib.ReferenceOrderStepper(o["st"], o["blocks"], o["dofs"].N, c["nu"], o["dt"])
    f = o["f"].copy()
    oracle_lib.collide_bgk(o["st"], f, c["nu"], o["dt"])
    for _ in range(5):
        stepper.step(f)
    # reference-ordered operators
    solver.collide()
    for _ in range(5):
        solver.stream()
        solver.collide()
    assert rel_err(solver.getF().to_host(), f) <= 1e-11
    assert rel_err(solver.getDensity(), stepper.rho) <= 1e-13
    # fused run() from the same start
    solver.setInitialFields(o["rho"], o["u"])
    solver.run(5)
    assert rel_err(solver.getF().to_host(), f) <= 1e-11
    solver.ctx.close()


# ---- entropic family ------------------------------------------------------------------------------------
ENTROPIC = [("D2Q9", "KBC_STANDARD"), ("D3Q15", "KBC_STANDARD"), ("D3Q19", "MRT_ENTROPIC")]


@pytest.mark.parametrize("in_init", [False, True])
@pytest.mark.parametrize("stencil,scheme", ENTROPIC)
def test_collide_entropic_matches_oracle(stencil, scheme, in_init, oracle_lib):
    """CollisionModel::collideAll of the legacy entropic models (KBCStandard.cpp:88-1028, MRTEntropic.cpp:167-305)
    on the BGKStandard_test population, tau_legacy = 0.9, dt = 0.1, scaled stencil."""
    from natrium_b200 import Context, Stencil, _capi
    scaling = 2.5
    st, ost = Stencil(stencil, scaling), oracle_lib.Stencil(stencil, scaling)
    n, dt = 1000, 0.1
    nu = 0.9 * dt * st.getSpeedOfSoundSquare()
    f = synthetic_populations(st.getQ(), n) * st.getWeights()[:, None]
    u0 = 0.05 * scaling * np.vstack([np.sin(np.arange(n) + d) for d in range(st.getD())])
    ctx = Context(0)
    ctx.set_stencil(st.getDirections(), st.getWeights(), st.getScaling(), st.getSpeedOfSoundSquare())
    ctx.set_layout(n, 0, False)
    ctx.set_dof_order(np.random.default_rng(5).permutation(n))
    ctx.set_collision(nu, dt, scheme=getattr(_capi, scheme), in_init=in_init)
    ctx.upload_populations(0, f)
    if in_init:
        ctx.upload_velocity(u0)
    ctx.collide()
    ctx.synchronize()
    got = ctx.download_populations(0)
    rho, u = ctx.download_moments()
    ref = f.copy()
    rrho, ru, rc = oracle_lib.collide_entropic(ost, ref, nu, dt, scheme, in_init=in_init, u_init=u0.copy() if in_init else None)
    assert rc == 0
    assert rel_err(got, ref) <= TOL_STEP
    assert rel_err(rho, rrho) <= 1e-14
    assert np.max(np.abs(u - ru)) <= 1e-13 * max(1.0, np.max(np.abs(ru)))
    if stencil != "D3Q15":      # KBCStandard D3Q15 mixes unscaled u^2 with the scaled cs2 (:749-750): mass drifts for scaling != 1
        assert np.max(np.abs(got.sum(axis=0) - f.sum(axis=0))) <= 1e-13 * np.max(f.sum(axis=0))
    ctx.close()


@pytest.mark.parametrize("case,scheme", [("c1_tgv2d_d2q9", "KBC_STANDARD"), ("tgv3d_d3q15", "KBC_STANDARD"),
                                         ("tgv3d_d3q19_small", "MRT_ENTROPIC")])
def test_fused_entropic_step_matches_oracle(case, scheme, oracle_lib):
    """nb200_step with an entropic collision == oracle stream (CSR vmult) followed by the oracle's collideAll, per step."""
    from natrium_b200 import _capi
    o = common.oracle_problem(case)
    ctx, c, st, pb, dt, part = make_ctx(case, fmt=("dict", 1e-14, True))
    ctx.set_collision(c["nu"], dt, scheme=getattr(_capi, scheme))
    f = o["f"].copy()
    ctx.upload_populations(0, f)
    rho_prev = None
    for s in range(5):
        ctx.step(1)
        ctx.synchronize()
        f = oracle_lib.stream(o["blocks"], f)
        rho_prev, _, rc = oracle_lib.collide_entropic(o["st"], f, c["nu"], dt, scheme, rho_prev=rho_prev)
        assert rc == 0
        got = ctx.download_populations(0)
        assert rel_err(got, f) <= TOL_STEP, (case, s, rel_err(got, f))
        ctx.upload_populations(0, f)
    assert rel_err(ctx.download_moments()[0], rho_prev) <= 1e-13
    ctx.close()


def test_entropic_dispatch_errors():
    """KBC_Standard only for D2Q9/D3Q15, MRT_ENTROPIC only for D3Q19: CollisionException otherwise."""
    from natrium_b200 import CollisionException, Context, Stencil, _capi
    for name, scheme in [("D3Q19", _capi.KBC_STANDARD), ("D3Q15", _capi.MRT_ENTROPIC), ("D2Q9", _capi.MRT_ENTROPIC)]:
        st = Stencil(name, 1.0)
        ctx = Context(0)
        ctx.set_stencil(st.getDirections(), st.getWeights(), 1.0, st.getSpeedOfSoundSquare())
        ctx.set_layout(8, 0, False)
        with pytest.raises(CollisionException):
            ctx.set_collision(0.1, 0.1, scheme=scheme)
        ctx.close()


# ---------------------------------------------------------------------------------------------
# collision_advanced Regularized / MultipleRelaxationTime / external forces
# ---------------------------------------------------------------------------------------------
ADVANCED = [("D2Q9", "BGK_REGULARIZED", None, None), ("D3Q15", "BGK_REGULARIZED", None, None),
            ("D3Q19", "BGK_REGULARIZED", None, None), ("D2Q9", "MRT_STANDARD", "DELLAR_D2Q9", "RELAX_FULL"),
            ("D2Q9", "MRT_STANDARD", "DELLAR_D2Q9", "DELLAR_RELAX_ONLY_N"), ("D2Q9", "MRT_STANDARD", "LALLEMAND_D2Q9", "RELAX_FULL"),
            ("D3Q19", "MRT_STANDARD", "DHUMIERES_D3Q19", "RELAX_FULL"), ("D3Q19", "MRT_STANDARD", "DHUMIERES_D3Q19", "RELAX_DHUMIERES_PAPER")]


def _set_advanced(ctx, st, nu, dt, scheme, basis, relax, **kw):
    from natrium_b200 import _capi, mrt
    if scheme == "MRT_STANDARD":
        tau = nu / (dt * st.getSpeedOfSoundSquare()) + 0.5
        b, r = getattr(mrt, basis), getattr(mrt, relax)
        ctx.set_mrt(mrt.make_M(b), mrt.make_T(b), mrt.make_diag(tau, b, r))
    ctx.set_collision(nu, dt, scheme=getattr(_capi, scheme), **kw)


@pytest.mark.parametrize("in_init", [False, True])
@pytest.mark.parametrize("stencil,scheme,basis,relax", ADVANCED)
def test_collide_advanced_matches_oracle(stencil, scheme, basis, relax, in_init, oracle_lib):
    """Regularized::relax / MultipleRelaxationTime::relax rows of selectCollision (CollisionSelection.h:87-88,182-186)
    on the BGKStandard_test population; the oracle uses the reference's own MRT literals, the product the host mirror."""
    from natrium_b200 import Context, Stencil
    scaling = 2.5
    st, ost = Stencil(stencil, scaling), oracle_lib.Stencil(stencil, scaling)
    n, dt = 1000, 0.1
    nu = 0.9 * dt * st.getSpeedOfSoundSquare()
    f = synthetic_populations(st.getQ(), n) * st.getWeights()[:, None]
    u0 = 0.05 * scaling * np.vstack([np.sin(np.arange(n) + d) for d in range(st.getD())])
    ctx = Context(0)
    ctx.set_stencil(st.getDirections(), st.getWeights(), st.getScaling(), st.getSpeedOfSoundSquare())
    ctx.set_layout(n, 0, False)
    _set_advanced(ctx, st, nu, dt, scheme, basis, relax, in_init=in_init)
    ctx.upload_populations(0, f)
    if in_init:
        ctx.upload_velocity(u0)
    ctx.collide()
    ctx.synchronize()
    got = ctx.download_populations(0)
    rho, u = ctx.download_moments()
    ref = f.copy()
    rrho, ru, rc = oracle_lib.collide_advanced(ost, ref, nu, dt, scheme=scheme, in_init=in_init,
                                               u_init=u0.copy() if in_init else None, mrt_basis=basis, relax_mode=relax or "RELAX_FULL")
    assert rc == 0
    assert rel_err(got, ref) <= TOL_STEP
    assert rel_err(rho, rrho) <= 1e-14
    assert np.max(np.abs(u - ru)) <= 1e-13 * max(1.0, np.max(np.abs(ru)))
    if not in_init:
        assert np.max(np.abs(got.sum(axis=0) - f.sum(axis=0))) <= 1e-13 * np.max(f.sum(axis=0))
    ctx.close()


@pytest.mark.parametrize("case,scheme,basis,relax", [("c1_tgv2d_d2q9", "BGK_REGULARIZED", None, None),
                                                     ("c1_tgv2d_d2q9", "MRT_STANDARD", "DELLAR_D2Q9", "RELAX_FULL"),
                                                     ("tgv3d_d3q15", "BGK_REGULARIZED", None, None),
                                                     ("tgv3d_d3q19_small", "BGK_REGULARIZED", None, None),
                                                     ("tgv3d_d3q19_small", "MRT_STANDARD", "DHUMIERES_D3Q19", "RELAX_FULL")])
def test_fused_advanced_step_matches_oracle(case, scheme, basis, relax, oracle_lib):
    """nb200_step with a Regularized / MRT collision == oracle stream followed by the oracle's collideAll, per step."""
    o = common.oracle_problem(case)
    ctx, c, st, pb, dt, part = make_ctx(case)
    _set_advanced(ctx, st, c["nu"], dt, scheme, basis, relax)
    f = o["f"].copy()
    ctx.upload_populations(0, f)
    for s in range(5):
        ctx.step(1)
        ctx.synchronize()
        f = oracle_lib.stream(o["blocks"], f)
        rrho, _, rc = oracle_lib.collide_advanced(o["st"], f, c["nu"], dt, scheme=scheme, mrt_basis=basis, relax_mode=relax or "RELAX_FULL")
        assert rc == 0
        got = ctx.download_populations(0)
        assert rel_err(got, f) <= TOL_STEP, (case, s, rel_err(got, f))
        ctx.upload_populations(0, f)
    assert rel_err(ctx.download_moments()[0], rrho) <= 1e-13
    ctx.close()


@pytest.mark.parametrize("in_init", [False, True])
@pytest.mark.parametrize("force_type", ["SHIFTING_VELOCITY", "EXACT_DIFFERENCE"])
@pytest.mark.parametrize("stencil,scheme,basis", [("D2Q9", "BGK_STANDARD", None), ("D3Q19", "BGK_STANDARD", None),
                                                  ("D3Q19", "BGK_REGULARIZED", None), ("D2Q9", "MRT_STANDARD", "LALLEMAND_D2Q9")])
def test_collide_forced_matches_oracle(stencil, scheme, basis, force_type, in_init, oracle_lib):
    """applyMacroscopicForces / applyForces / postCollisionApplyForces (AuxiliaryCollisionFunctions.h:332-417) inside
    collideAll, f only."""
    from natrium_b200 import Context, Stencil, _capi
    scaling = 2.0
    st, ost = Stencil(stencil, scaling), oracle_lib.Stencil(stencil, scaling)
    n, dt = 640, 0.1
    nu = 0.9 * dt * st.getSpeedOfSoundSquare()
    F = np.array([1e-2, -2e-2, 5e-3])[:st.getD()]
    f = synthetic_populations(st.getQ(), n) * st.getWeights()[:, None]
    u0 = 0.05 * scaling * np.vstack([np.sin(np.arange(n) + d) for d in range(st.getD())])
    ctx = Context(0)
    ctx.set_stencil(st.getDirections(), st.getWeights(), st.getScaling(), st.getSpeedOfSoundSquare())
    ctx.set_layout(n, 0, False)
    _set_advanced(ctx, st, nu, dt, scheme, basis, "RELAX_FULL", in_init=in_init, force=F, force_type=getattr(_capi, force_type))
    ctx.upload_populations(0, f)
    if in_init:
        ctx.upload_velocity(u0)
    ctx.collide()
    ctx.synchronize()
    got = ctx.download_populations(0)
    rho, u = ctx.download_moments()
    ref = f.copy()
    rrho, ru, rc = oracle_lib.collide_advanced(ost, ref, nu, dt, scheme=scheme, in_init=in_init, u_init=u0.copy() if in_init else None,
                                               force=F, force_type=force_type, mrt_basis=basis)
    assert rc == 0
    assert rel_err(got, ref) <= TOL_STEP
    assert rel_err(rho, rrho) <= 1e-14
    assert np.max(np.abs(u - ru)) <= 1e-13 * max(1.0, np.max(np.abs(ru)))
    ctx.close()


@pytest.mark.parametrize("force_type", ["SHIFTING_VELOCITY", "EXACT_DIFFERENCE"])
@pytest.mark.parametrize("stencil", ["D2Q25H", "D3Q45"])
def test_collide_fg_forced_matches_oracle(stencil, force_type, oracle_lib):
    """f + g collideAll with an external force (the channel configuration uses EXACT_DIFFERENCE, step-turbulent-channel.cpp)."""
    from natrium_b200 import Context, Stencil, _capi, harness
    st, ost = Stencil(stencil, 1.0), oracle_lib.Stencil(stencil, 1.0)
    n, dt, nu, gamma = 500, 0.05, 0.002, 1.4
    rng = np.random.default_rng(11)
    rho = 1.0 + 0.1 * rng.standard_normal(n)
    u = 0.1 * rng.standard_normal((st.getD(), n))
    T = 1.0 + 0.05 * rng.standard_normal(n)
    F = np.array([3e-2, 0.0, -1e-2])[:st.getD()]
    f, g = harness.quartic_equilibrium_distributions(st, rho, u, T, gamma)
    f *= 1.0 + 0.01 * rng.standard_normal(f.shape)
    g *= 1.0 + 0.01 * rng.standard_normal(g.shape)
    ctx = Context(0)
    ctx.set_stencil(st.getDirections(), st.getWeights(), st.getScaling(), st.getSpeedOfSoundSquare())
    ctx.set_layout(n, 0, True)
    ctx.set_collision(nu, dt, equilibrium=_capi.QUARTIC_EQUILIBRIUM, with_g=True, gamma=gamma, prandtl=0.7, sutherland=True,
                      force=F, force_type=getattr(_capi, force_type))
    ctx.upload_populations(0, f)
    ctx.upload_populations(1, g)
    ctx.collide()
    ctx.synchronize()
    gf, gg = ctx.download_populations(0), ctx.download_populations(1)
    grho, gu, gT, gs = ctx.download_moments(want_T=True)
    rf, rg = f.copy(), g.copy()
    rrho, ru, rT, rs, rc = oracle_lib.collide_bgk_fg_forced(ost, rf, rg, nu, dt, F, force_type, gamma=gamma, prandtl=0.7, sutherland=True)
    assert rc == 0
    assert rel_err(gf, rf) <= TOL_STEP and rel_err(gg, rg) <= TOL_STEP
    assert rel_err(grho, rrho) <= 1e-14 and rel_err(gT, rT) <= 1e-12
    assert np.max(np.abs(gu - ru)) <= 1e-13 * max(1.0, np.max(np.abs(ru)))
    ctx.close()


def test_forced_step_runs_unfused_and_matches_oracle(oracle_lib):
    """nb200_step with an external force: stream + collide as two kernels, same result as the oracle's order."""
    from natrium_b200 import _capi
    case = "c1_tgv2d_d2q9"
    o = common.oracle_problem(case)
    ctx, c, st, pb, dt, part = make_ctx(case)
    F = np.array([0.5, -0.25])
    ctx.set_collision(c["nu"], dt, force=F, force_type=_capi.EXACT_DIFFERENCE)
    f = o["f"].copy()
    ctx.upload_populations(0, f)
    l0 = ctx.kernel_launches()
    ctx.step(3)
    ctx.synchronize()
    assert ctx.kernel_launches() - l0 == 6
    for _ in range(3):
        f = oracle_lib.stream(o["blocks"], f)
        _, _, rc = oracle_lib.collide_advanced(o["st"], f, c["nu"], dt, force=F, force_type="EXACT_DIFFERENCE")
        assert rc == 0
    assert rel_err(ctx.download_populations(0), f) <= 3 * TOL_STEP
    ctx.close()


def test_advanced_dispatch_errors():
    """selectCollision rows that do not exist throw; forcing switched off with a force present throws
    (NATriuMException, Aux...h:335-339); GUO is "not implemented"; MRT needs its tables."""
    from natrium_b200 import CollisionException, Context, NatriumB200Error, Stencil, _capi, mrt
    def ctx_for(name, with_g=False):
        st = Stencil(name, 1.0)
        ctx = Context(0)
        ctx.set_stencil(st.getDirections(), st.getWeights(), 1.0, st.getSpeedOfSoundSquare())
        ctx.set_layout(8, 0, with_g)
        return ctx
    for name, scheme in [("D3Q15", _capi.MRT_STANDARD), ("D2Q25H", _capi.BGK_REGULARIZED), ("D3Q45", _capi.MRT_STANDARD)]:
        ctx = ctx_for(name)
        with pytest.raises(CollisionException):
            ctx.set_collision(0.1, 0.1, scheme=scheme)
        ctx.close()
    ctx = ctx_for("D3Q19")
    with pytest.raises(NatriumB200Error):          # tables missing
        ctx.set_collision(0.1, 0.1, scheme=_capi.MRT_STANDARD)
    with pytest.raises(NatriumB200Error):          # wrong size
        ctx.set_mrt(mrt.make_M(mrt.DELLAR_D2Q9), mrt.make_T(mrt.DELLAR_D2Q9), mrt.make_diag(0.8, mrt.DELLAR_D2Q9))
    with pytest.raises(NatriumB200Error) as ei:
        ctx.set_collision(0.1, 0.1, force=[1e-3, 0, 0], force_type=_capi.NO_FORCING)
    assert "forcing was switched off" in str(ei.value)
    with pytest.raises(CollisionException) as ei:
        ctx.set_collision(0.1, 0.1, force=[1e-3, 0, 0], force_type=_capi.GUO)
    assert "Force Type not implemented" in str(ei.value)
    with pytest.raises(CollisionException):
        ctx.set_collision(0.1, 0.1, scheme=_capi.MRT_ENTROPIC, force=[1e-3, 0, 0], force_type=_capi.SHIFTING_VELOCITY)
    ctx.close()
    # reference quirk: compressible D2Q25H with BGK_REGULARIZED / BGK_EQUILIBRIUM runs the plain BGK collision
    ctx = ctx_for("D2Q25H", with_g=True)
    ctx.set_collision(0.1, 0.1, scheme=_capi.BGK_REGULARIZED, equilibrium=_capi.BGK_EQUILIBRIUM, with_g=True)
    ctx.close()


# ---------------------------------------------------------------------------------------------
# wall hits
# ---------------------------------------------------------------------------------------------
def _synthetic_hits(rng, n, Q, n_hits, thermal):
    """A flattened HitList: destination DoFs with several hits each (corner nodes are hit from more than one
    direction), in arbitrary cell order."""
    idx = rng.integers(0, n, size=n_hits).astype(np.int32)
    idx[1::3] = idx[0:-1:3][:len(idx[1::3])]                 # repeated destinations
    dirs = rng.integers(1, Q, size=n_hits).astype(np.int32)
    kinds = (rng.random(n_hits) < 0.5).astype(np.int32) if thermal else np.zeros(n_hits, dtype=np.int32)
    vals = np.where(kinds == 1, 0.85, 1e-2 * rng.standard_normal(n_hits))
    return idx, dirs, kinds, vals


@pytest.mark.parametrize("ordered", [False, True])
def test_stream_with_wall_hits_f_only(ordered, oracle_lib):
    """SemiLagrangian::stream = vmult + boundaryHandler.apply (SemiLagrangian.h:150-161): VelocityNeqBounceBack terms on
    a D2Q9 problem, with and without an internal DoF order."""
    case = "c1_tgv2d_d2q9"
    o = common.oracle_problem(case)
    ctx, c, st, pb, dt, part = make_ctx(case, fmt=("dict", 1e-14, ordered))
    rng = np.random.default_rng(2)
    idx, dirs, kinds, vals = _synthetic_hits(rng, part.n_owned, 9, 300, thermal=False)
    ctx.set_wall_hits(idx, dirs, kinds, vals)
    ctx.upload_populations(0, o["f"])
    ctx.stream(0)
    got = ctx.download_populations(0)
    ref = oracle_lib.stream(o["blocks"], o["f"])
    assert oracle_lib.apply_wall_hits(o["st"], ref, None, idx, dirs, kinds, vals) == 0
    assert rel_err(got, ref) <= TOL_STEP
    # fused step path falls back to stream -> hits -> collide
    ctx.set_collision(c["nu"], dt)
    ctx.upload_populations(0, o["f"])
    ctx.step(2)
    ctx.synchronize()
    f = o["f"].copy()
    for _ in range(2):
        f = oracle_lib.stream(o["blocks"], f)
        oracle_lib.apply_wall_hits(o["st"], f, None, idx, dirs, kinds, vals)
        oracle_lib.collide_bgk(o["st"], f, c["nu"], dt)
    assert rel_err(ctx.download_populations(0), f) <= 2 * TOL_STEP
    ctx.set_wall_hits([], [], [], [])        # clearing the hit list restores the fused path
    l0 = ctx.kernel_launches()
    ctx.step(1)
    assert ctx.kernel_launches() - l0 == 1
    ctx.close()


def test_compressible_step_with_thermal_walls(oracle_lib):
    """CompressibleCFDSolver order (CompressibleCFDSolver.h:181-314): stream f, wall hits on the new f and the old g
    (ThermalBounceBack re-equilibration + VelocityNeqBounceBack terms), then gStream, then collide."""
    from natrium_b200 import _capi
    case = "tgv3d_d3q45"
    o = common.oracle_problem(case)
    ctx, c, st, pb, dt, part = make_ctx(case)
    set_collision(ctx, c, dt)
    rng = np.random.default_rng(4)
    idx, dirs, kinds, vals = _synthetic_hits(rng, part.n_owned, 45, 200, thermal=True)
    ctx.set_wall_hits(idx, dirs, kinds, vals)
    f, g = o["f"].copy(), o["g"].copy()
    ctx.upload_populations(0, f)
    ctx.upload_populations(1, g)
    for s in range(3):
        ctx.step(1)
        ctx.synchronize()
        f = oracle_lib.stream(o["blocks"], f)
        assert oracle_lib.apply_wall_hits(o["st"], f, g, idx, dirs, kinds, vals) == 0
        g = oracle_lib.stream(o["blocks"], g)
        _, _, _, _, rc = oracle_lib.collide_bgk_fg(o["st"], f, g, c["nu"], dt, equilibrium=1, gamma=1.4, prandtl=0.71, sutherland=True)
        assert rc == 0
        gf, gg = ctx.download_populations(0), ctx.download_populations(1)
        assert rel_err(gf, f) <= TOL_STEP and rel_err(gg, g) <= TOL_STEP, (s, rel_err(gf, f), rel_err(gg, g))
        ctx.upload_populations(0, f)
        ctx.upload_populations(1, g)
    ctx.close()


def test_wall_hit_errors():
    from natrium_b200 import CollisionException, Context, NatriumB200Error, Stencil, _capi
    st = Stencil("D3Q19", 1.0)
    ctx = Context(0)
    ctx.set_stencil(st.getDirections(), st.getWeights(), 1.0, st.getSpeedOfSoundSquare())
    ctx.set_layout(16, 0, False)
    with pytest.raises(CollisionException):      # thermal walls are a D3Q45 f+g model
        ctx.set_wall_hits([1], [2], [_capi.WALL_THERMAL_BOUNCE_BACK], [0.85])
    with pytest.raises(NatriumB200Error):
        ctx.set_wall_hits([16], [2], [0], [0.1])
    with pytest.raises(NatriumB200Error):
        ctx.set_wall_hits([1], [19], [0], [0.1])
    with pytest.raises(CollisionException):
        ctx.set_wall_hits([1], [2], [7], [0.1])
    ctx.close()


# ---------------------------------------------------------------------------------------------
# host-buffer step (nb200_step_host)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case,ordered,chunks", [("c1_tgv2d_d2q9", False, 2), ("c1_tgv2d_d2q9", True, 2), ("tgv3d_d3q19_small", True, 4),
                                                 ("tgv3d_d3q19_small", False, 3), ("tgv3d_d3q19_small", True, 1), ("tgv2d_d2q25", True, 4)])
def test_step_host_equals_device_resident_step(case, ordered, chunks, oracle_lib):
    """The chunk-pipelined host-buffer step gives bit-for-bit what upload + nb200_step + download gives (the kernels
    and their per-row arithmetic are the same; only the launch is cut into pieces), over several chained steps with
    the output buffer of one step feeding the next; f+g and chunks = 1 take the sequential legs."""
    import torch
    o = common.oracle_problem(case)
    ctx, c, st, pb, dt, part = make_ctx(case, fmt=("dict", 1e-14, ordered))
    set_collision(ctx, c, dt)
    n, Q, D = part.n_owned, st.getQ(), st.getD()
    with_g = bool(c.get("with_g"))
    # reference: device-resident steps
    ctx.upload_populations(0, o["f"])
    if with_g:
        ctx.upload_populations(1, o["g"])
    ctx.step(3)
    ctx.synchronize()
    want = ctx.download_populations(0)
    wrho, wu = ctx.download_moments()[:2]
    # host-buffer steps
    a = torch.empty((Q, n), dtype=torch.float64, pin_memory=True)
    b = torch.empty((Q, n), dtype=torch.float64, pin_memory=True)
    mom = torch.empty((1 + D, n), dtype=torch.float64, pin_memory=True)
    a.numpy()[...] = o["f"]
    if with_g:
        ctx.upload_populations(1, o["g"])
    bufs = [a, b]
    l0 = ctx.kernel_launches()
    for s in range(3):
        ctx.step_host(bufs[s & 1].data_ptr(), bufs[(s + 1) & 1].data_ptr(), mom.data_ptr(), mom.data_ptr() + 8 * n, chunks)
    ctx.synchronize()
    got = bufs[3 & 1].numpy()
    if with_g:
        assert rel_err(got, want) <= TOL_STEP       # g stays on the device between the host steps
    else:
        assert np.array_equal(got, want)
        assert np.array_equal(mom.numpy()[0], wrho) and np.array_equal(mom.numpy()[1:], wu)
    if chunks > 1 and not with_g:
        # per piece: the fused kernel, plus scatter / gather kernels when the library permutes the host numbering
        assert ctx.kernel_launches() - l0 >= 3 * chunks * (3 if ordered else 1)
    ctx.close()


# ---------------------------------------------------------------------------------------------
# pseudo-entropic stabilizer
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case,with_e", [("c1_tgv2d_d2q9", False), ("c1_tgv2d_d2q9", True), ("tgv3d_d3q19_small", False)])
def test_step_with_stabilizer_matches_oracle(case, with_e, oracle_lib):
    """run() loop body with the PseudoEntropicStabilizer data processor (CFDSolver.cpp:884-891): stream, collide, then
    f <- A f.  The oracle multiplies by the reference's literal matrix, the product by the host mirror's."""
    from natrium_b200 import mrt
    o = common.oracle_problem(case)
    ctx, c, st, pb, dt, part = make_ctx(case)
    set_collision(ctx, c, dt)
    name = c["stencil"]
    ctx.set_post_collision_matrix(mrt.make_stabilizer(name, with_e))
    A = oracle_lib.stabilizer_matrix(name, with_e)
    stepper = oracle_lib.ReferenceOrderStepper(o["st"], o["blocks"], o["dofs"].N, c["nu"], dt)
    f = o["f"].copy()
    ctx.upload_populations(0, f)
    for s in range(5):
        ctx.step(1)
        ctx.synchronize()
        stepper.step(f)
        oracle_lib.apply_stabilizer(f, A)
        got = ctx.download_populations(0)
        assert rel_err(got, f) <= TOL_STEP, (s, rel_err(got, f))
        ctx.upload_populations(0, f)
    # explicit application == DataProcessor::apply()
    ctx.apply_post_collision()
    oracle_lib.apply_stabilizer(f, A)
    assert rel_err(ctx.download_populations(0), f) <= TOL_STEP
    ctx.set_post_collision_matrix(None)
    l0 = ctx.kernel_launches()
    ctx.step(1)
    assert ctx.kernel_launches() - l0 == 1
    ctx.close()


def test_config1_integration_with_stabilizer(oracle_lib):
    """Integration test #11 (ConvergenceTestSemiLagrangianPeriodic, IntegrationTestCases.cpp:885-961) through the host
    mirror: CFDSolver + appended PseudoEntropicStabilizer, device-resident run to t = 1/(2 nu);
    E_kin(t)/E_kin(0) = exp(-2) +- 1e-2."""
    from natrium_b200 import PseudoEntropicStabilizer, Stencil
    from natrium_b200 import Context, harness
    case = "c1_tgv2d_d2q9"
    o = common.oracle_problem(case)
    ctx, c, st, pb, dt, part = make_ctx(case)
    set_collision(ctx, c, dt)
    from natrium_b200 import mrt
    ctx.set_post_collision_matrix(mrt.make_stabilizer("D2Q9"))
    ctx.upload_populations(0, o["f"])
    ctx.collide()
    E0 = ctx.conserved()[4]
    steps = int(round(0.5 / dt))
    ctx.step(steps)
    ctx.synchronize()
    ratio = ctx.conserved()[4] / E0
    assert abs(ratio - np.exp(-4 * c["nu"] * steps * dt)) < 1e-2, ratio
    with pytest.raises(Exception):
        ctx45, *_ = make_ctx("tgv3d_d3q45", with_matrix=False)
        try:
            ctx45.set_post_collision_matrix(np.eye(45))
        finally:
            ctx45.close()
    ctx.close()


# ---------------------------------------------------------------------------------------------
# walled, stretched meshes assembled by the oracle (C4 / C5-like): bounce blocks + hit list + collide
# ---------------------------------------------------------------------------------------------
def _opposite(e):
    return np.array([int(np.argmin(np.abs(e + e[i]).sum(1))) for i in range(len(e))])


def _walled_problem(oracle_lib, name, scaling, verts, boundary, p, cfl):
    from oracle import assembly
    ost = oracle_lib.Stencil(name, scaling)
    mesh = assembly.CartesianMesh([np.asarray(v, dtype=np.float64) for v in verts], boundary=boundary)
    dt = assembly.calculate_timestep(mesh, p, ost.max_speed, cfl)
    blocks, dofs = assembly.assemble_semilagrangian(mesh, p, ost.e, dt, opposite=_opposite(ost.e))
    return ost, mesh, dt, blocks, dofs


def dofmap_grid_coords(dofs):
    """(dims, coords) for nb200_set_dof_grid from the oracle's DofMap (lexicographic tensor grid, optional renumbering)."""
    nd = list(dofs.nd)
    i = np.arange(dofs.N)
    lexc = np.stack([i % nd[0], i // nd[0]] if len(nd) == 2 else [i % nd[0], (i // nd[0]) % nd[1], i // (nd[0] * nd[1])], axis=1)
    if dofs.numbering is None:
        return nd, lexc
    coords = np.empty_like(lexc)
    coords[dofs.numbering] = lexc
    return nd, coords


def _upload_oracle_blocks(ctx, blocks):
    for (bi, bj), m in sorted(blocks.items()):
        ctx.upload_block_csr(bi, bj, m.indptr, m.indices, m.data)
    ctx.finalize_matrix()


@pytest.mark.parametrize("fmt", [None, ("dict", 1e-14, True), ("dict-unstaged", 0.0, False), ("ell", 0.0, False), ("dict", 1e-14, False, "grid")],
                         ids=["dict-default", "dict-permuted", "dict-unstaged", "ell", "grid"])
def test_lid_driven_walled_stretched_2d(fmt, oracle_lib):
    """C4-like: D2Q9 on a y-stretched mesh with VelocityNeqBounceBack walls in y (moving lid) and periodic x.  The
    matrix (off-diagonal bounce blocks, SemiLagrangian.cpp:358-384) and the hit list (addHit, :381-383) come from the
    oracle's restatement of fillSparseObject; the per-hit term 2 w rho e.u_wall/cs2 is evaluated on the host as the
    reference does (VelocityNeqBounceBack.cpp:137-195).  Five steps: stream -> wall hits -> collide."""
    from natrium_b200 import Context, Stencil, _capi
    name, scaling, p = "D2Q9", 3.0, 3
    ost, mesh, dt, blocks, dofs = _walled_problem(oracle_lib, name, scaling, [np.linspace(0, 2.0, 6), 2.0 * np.array([0, 0.2, 0.45, 0.75, 1.0])],
                                                  ["periodic", "wall"], p, 0.8)
    assert any(bi != bj for bi, bj in blocks) and len(dofs.hits) > 0
    st = Stencil(name, scaling)
    n = dofs.N
    u_lid = np.array([0.3, 0.0])
    hits = dofs.hits
    idx = np.array([h["index"] for h in hits], dtype=np.int32)
    dirs = np.array([h["direction"] for h in hits], dtype=np.int32)
    vals = np.array([2 * ost.w[h["direction"]] * 1.0 * float(ost.e[h["direction"]] @ (u_lid if h["boundary"] == (1, 1) else 0 * u_lid)) / ost.cs2
                     for h in hits])
    kinds = np.zeros(len(hits), dtype=np.int32)
    ctx = Context(0)
    ctx.set_stencil(st.getDirections(), st.getWeights(), st.getScaling(), st.getSpeedOfSoundSquare())
    ctx.set_layout(n, 0, False)
    if fmt is not None:
        ctx.set_matrix_format(_fmt_code(fmt[0]), fmt[1])
        if fmt[2]:
            ctx.set_dof_order(np.random.default_rng(9).permutation(n))
        if len(fmt) > 3:
            ctx.set_dof_grid(*dofmap_grid_coords(dofs), fe_order=p)
    _upload_oracle_blocks(ctx, blocks)
    ctx.set_wall_hits(idx, dirs, kinds, vals)
    nu = 0.05
    ctx.set_collision(nu, dt)
    x = dofs.support_points()
    rho0 = 1.0 + 0.02 * np.cos(np.pi * x[:, 0])
    u0 = np.stack([0.05 * np.sin(np.pi * x[:, 1]), 0.02 * np.cos(np.pi * x[:, 0])])
    from oracle import fields
    f = fields.equilibrium_init(ost.e, ost.w, ost.cs2, rho0, u0)
    ctx.upload_populations(0, f)
    for s in range(5):
        ctx.step(1)
        ctx.synchronize()
        f = oracle_lib.stream(blocks, f)
        assert oracle_lib.apply_wall_hits(ost, f, None, idx, dirs, kinds, vals) == 0
        _, _, rc = oracle_lib.collide_bgk(ost, f, nu, dt)
        assert rc == 0
        got = ctx.download_populations(0)
        assert rel_err(got, f) <= TOL_STEP, (s, rel_err(got, f))
        ctx.upload_populations(0, f)
    ctx.close()


def test_thermal_channel_walled_stretched_3d(oracle_lib):
    """C5-like: D3Q45 f+g on a y-stretched mesh (TurbulentChannelFlow3D-style grading), ThermalBounceBack walls in y
    (T_w = 0.85), periodic x and z, EXACT_DIFFERENCE forcing, Pr = 0.7, Sutherland law: stream f -> wall hits (f new,
    g old) -> gStream -> forced collide, against the oracle in reference order."""
    from natrium_b200 import Context, Stencil, _capi, harness
    name, p = "D3Q45", 2
    y = np.linspace(0, 1, 4)
    vy = 2.0 * (y - 0.8 * np.sin(2 * np.pi * y) / (2 * np.pi))          # TurbulentChannelFlow3D.h:116-123
    ost, mesh, dt, blocks, dofs = _walled_problem(oracle_lib, name, 1.0, [np.linspace(0, 2.0, 3), vy, np.linspace(0, 1.0, 2)],
                                                  ["periodic", "wall", "periodic"], p, 0.4)
    st = Stencil(name, 1.0)
    n = dofs.N
    hits = dofs.hits
    assert len(hits) > 0
    idx = np.array([h["index"] for h in hits], dtype=np.int32)
    dirs = np.array([h["direction"] for h in hits], dtype=np.int32)
    kinds = np.full(len(hits), _capi.WALL_THERMAL_BOUNCE_BACK, dtype=np.int32)
    vals = np.full(len(hits), 0.85)
    ctx = Context(0)
    ctx.set_stencil(st.getDirections(), st.getWeights(), st.getScaling(), st.getSpeedOfSoundSquare())
    ctx.set_layout(n, 0, True)
    _upload_oracle_blocks(ctx, blocks)
    ctx.set_wall_hits(idx, dirs, kinds, vals)
    nu, gamma, F = 0.01, 1.4, np.array([2e-2, 0.0, 0.0])
    ctx.set_collision(nu, dt, equilibrium=_capi.QUARTIC_EQUILIBRIUM, with_g=True, gamma=gamma, prandtl=0.7, sutherland=True,
                      force=F, force_type=_capi.EXACT_DIFFERENCE)
    x = dofs.support_points()
    rho0 = 1.0 + 0.02 * np.cos(np.pi * x[:, 0])
    u0 = np.stack([0.05 * np.sin(np.pi * x[:, 1] / 2.0), 0.0 * x[:, 0], 0.01 * np.cos(np.pi * x[:, 0])])
    T0 = 0.85 + 0.1 * np.sin(np.pi * x[:, 1] / 2.0)
    f, g = harness.quartic_equilibrium_distributions(st, rho0, u0, T0, gamma)
    ctx.upload_populations(0, f)
    ctx.upload_populations(1, g)
    for s in range(3):
        ctx.step(1)
        ctx.synchronize()
        f = oracle_lib.stream(blocks, f)
        assert oracle_lib.apply_wall_hits(ost, f, g, idx, dirs, kinds, vals) == 0
        g = oracle_lib.stream(blocks, g)
        _, _, _, _, rc = oracle_lib.collide_bgk_fg_forced(ost, f, g, nu, dt, F, "EXACT_DIFFERENCE", gamma=gamma, prandtl=0.7, sutherland=True)
        assert rc == 0
        gf, gg = ctx.download_populations(0), ctx.download_populations(1)
        assert rel_err(gf, f) <= TOL_STEP and rel_err(gg, g) <= TOL_STEP, (s, rel_err(gf, f), rel_err(gg, g))
        ctx.upload_populations(0, f)
        ctx.upload_populations(1, g)
    ctx.close()


def test_cell_numbering_equals_order_hint(oracle_lib):
    """A host that numbers its DoFs cell by cell (deal.II) and passes no hint gets bit for bit what a lexicographic host
    with the nb200_set_dof_order hint gets, and nb200_step_host then copies straight between host buffers and device
    arrays (no permutation kernels: 1 launch per pipelined piece)."""
    import torch
    from natrium_b200 import Context, harness
    case = "tgv3d_d3q19_small"
    o = common.oracle_problem(case)
    ctx_a, c, st, pb, dt, part = make_ctx(case, fmt=("dict", 1e-14, True))
    set_collision(ctx_a, c, dt)
    ctx_a.upload_populations(0, o["f"])
    ctx_a.step(4)
    ctx_a.synchronize()
    want = ctx_a.download_populations(0)
    ctx_a.close()
    num = harness.CellNumbering(part)
    ctx = Context(0)
    ctx.set_stencil(st.getDirections(), st.getWeights(), st.getScaling(), st.getSpeedOfSoundSquare())
    ctx.set_layout(part.n_owned, part.n_ghost, False)
    harness.upload_streaming_matrix(ctx, pb, part, st, dt, numbering=num)
    assert ctx.matrix_format_info()["staged"]
    set_collision(ctx, c, dt)
    ctx.upload_populations(0, np.ascontiguousarray(o["f"][:, num.order]))
    ctx.step(4)
    ctx.synchronize()
    got = ctx.download_populations(0)
    assert np.array_equal(got, want[:, num.order])
    n, Q, D = part.n_owned, st.getQ(), st.getD()
    a = torch.empty((Q, n), dtype=torch.float64, pin_memory=True)
    b = torch.empty((Q, n), dtype=torch.float64, pin_memory=True)
    mom = torch.empty((1 + D, n), dtype=torch.float64, pin_memory=True)
    a.numpy()[...] = o["f"][:, num.order]
    l0 = ctx.kernel_launches()
    bufs = [a, b]
    for s in range(4):
        ctx.step_host(bufs[s & 1].data_ptr(), bufs[(s + 1) & 1].data_ptr(), mom.data_ptr(), mom.data_ptr() + 8 * n, 4)
    ctx.synchronize()
    assert np.array_equal(bufs[0].numpy(), want[:, num.order])
    assert ctx.kernel_launches() - l0 == 4 * 4
    ctx.close()


# ---------------------------------------------------------------------------------------------
# degenerate sizes: a rank that owns nothing, and DoF counts around the warp / CTA boundaries
# ---------------------------------------------------------------------------------------------
def test_rank_without_dofs():
    """More ranks than cell layers leaves ranks without owned DoFs in a p4est partition: every call must be a no-op."""
    import scipy.sparse as sp
    from natrium_b200 import Context, Stencil
    st = Stencil("D2Q9", 1.0)
    ctx = Context(0)
    ctx.set_stencil(st.getDirections(), st.getWeights(), 1.0, st.getSpeedOfSoundSquare())
    ctx.set_layout(0, 0, False)
    empty = sp.csr_matrix((0, 0))
    for a in range(8):
        ctx.upload_block_csr(a, a, np.zeros(1, dtype=np.int64), np.zeros(0, dtype=np.int32), np.zeros(0))
    ctx.finalize_matrix()
    ctx.set_collision(0.1, 0.1)
    ctx.upload_populations(0, np.zeros((9, 0)))
    ctx.stream(0)
    ctx.collide()
    ctx.step(3)
    ctx.synchronize()
    assert ctx.download_populations(0).shape == (9, 0)
    assert np.all(ctx.conserved() == 0.0)
    ctx.close()


@pytest.mark.parametrize("fmt", [None, ("dict-unstaged", 1e-14, False), ("ell", 0.0, False)], ids=["dict", "dict-unstaged", "ell"])
@pytest.mark.parametrize("n", [1, 31, 33, 127, 129, 257])
def test_sizes_around_warp_and_cta_boundaries(n, fmt, oracle_lib):
    """Ragged tails: n_owned not a multiple of 32 / 64 / 128 (the row pairing works on 64-row halves, the ELL on 32-row
    slices).  Random sparse periodic shift matrices so that every row reads its neighbours; one fused step and one
    stream against the oracle."""
    import scipy.sparse as sp
    from natrium_b200 import Context, Stencil
    st = Stencil("D2Q9", 1.0)
    ost = oracle_lib.Stencil("D2Q9", 1.0)
    rng = np.random.default_rng(n)
    blocks = {}
    for a in range(8):
        k = min(n, 3)
        cols = (np.arange(n)[:, None] + rng.integers(0, n, size=(1, k))) % n
        vals = rng.random((n, k)) + 0.1
        vals /= vals.sum(1, keepdims=True)
        m = sp.csr_matrix((vals.reshape(-1), cols.reshape(-1), np.arange(n + 1) * k), shape=(n, n))
        m.sum_duplicates()
        m.sort_indices()
        blocks[(a, a)] = m
    f = (1.0 + 0.1 * rng.random((9, n))) * st.getWeights()[:, None]
    ctx = Context(0)
    ctx.set_stencil(st.getDirections(), st.getWeights(), 1.0, st.getSpeedOfSoundSquare())
    ctx.set_layout(n, 0, False)
    if fmt is not None:
        ctx.set_matrix_format(_fmt_code(fmt[0]), fmt[1])
    for (bi, bj), m in blocks.items():
        ctx.upload_block_csr(bi, bj, m.indptr, m.indices, m.data)
    ctx.finalize_matrix()
    ctx.set_collision(0.05, 0.1)
    ctx.upload_populations(0, f)
    ctx.stream(0)
    ref = oracle_lib.stream(blocks, f)
    assert rel_err(ctx.download_populations(0), ref) <= TOL_STEP
    ctx.upload_populations(0, f)
    ctx.step(2)
    ctx.synchronize()
    g = f.copy()
    for _ in range(2):
        g = oracle_lib.stream(blocks, g)
        oracle_lib.collide_bgk(ost, g, 0.05, 0.1)
    assert rel_err(ctx.download_populations(0), g) <= 2 * TOL_STEP
    ctx.close()


def test_collision_selection_table():
    """CollisionSelection2D/3D_test (test/collision_advanced/CollisionSelection_test.cpp:33-214) for the stencils on the
    path: every row of selectCollision (CollisionSelection.h:85-91,141-150,179-202,251) is accepted, every other
    combination throws "Collision model not implemented yet"."""
    from natrium_b200 import CollisionException, Context, Stencil, _capi, mrt
    B, R, M = _capi.BGK_STANDARD, _capi.BGK_REGULARIZED, _capi.MRT_STANDARD
    E, Qe = _capi.BGK_EQUILIBRIUM, _capi.QUARTIC_EQUILIBRIUM
    rows_f = {"D2Q9": {(B, E), (B, Qe), (R, E), (M, E)}, "D2Q25H": {(B, Qe)}, "D3Q15": {(B, E), (R, E)},
              "D3Q19": {(B, E), (R, E), (M, E)}, "D3Q45": {(B, E), (B, Qe)}}
    rows_fg = {"D2Q25H": {(B, E), (B, Qe), (R, E)}, "D3Q45": {(B, Qe)}}
    for with_g, table in ((False, rows_f), (True, rows_fg)):
        for name in ("D2Q9", "D2Q25H", "D3Q15", "D3Q19", "D3Q45"):
            st = Stencil(name, 1.0)
            ctx = Context(0)
            ctx.set_stencil(st.getDirections(), st.getWeights(), 1.0, st.getSpeedOfSoundSquare())
            ctx.set_layout(4, 0, with_g)
            if st.getQ() in (9, 19):
                b = mrt.DELLAR_D2Q9 if st.getQ() == 9 else mrt.DHUMIERES_D3Q19
                ctx.set_mrt(mrt.make_M(b), mrt.make_T(b), mrt.make_diag(0.8, b))
            for scheme in (B, R, M):
                for eq in (E, Qe):
                    ok = (scheme, eq) in table.get(name, set())
                    try:
                        ctx.set_collision(0.1, 0.1, scheme=scheme, equilibrium=eq, with_g=with_g)
                        assert ok, (name, scheme, eq, with_g, "accepted but not a row of selectCollision")
                    except CollisionException as ex:
                        assert not ok, (name, scheme, eq, with_g, str(ex))
                        assert "not implemented" in str(ex)
            ctx.close()


def test_poiseuille_bounce_back_with_forcing_on_device(oracle_lib):
    """SemiLagrangianBoundaryHandler_PoiseuilleBB_test (test/boundaries/SemiLagrangianBoundaryHandler_test.cpp:175-229) on
    the device: bounce blocks + wall hits + SHIFTING_VELOCITY forcing in nb200_step.  The trajectory follows the oracle
    (200 steps), and the converged integral mean of u_x lies within 10 % of u_bulk like the reference demands."""
    from natrium_b200 import Context, Stencil, _capi
    pb = common.poiseuille_problem(oracle_lib)
    ost, n = pb["st"], pb["dofs"].N
    st = Stencil("D2Q9", pb["scaling"])
    idx, dirs, kinds, vals = pb["hits"]
    ctx = Context(0)
    ctx.set_stencil(st.getDirections(), st.getWeights(), st.getScaling(), st.getSpeedOfSoundSquare())
    ctx.set_layout(n, 0, False)
    _upload_oracle_blocks(ctx, pb["blocks"])
    ctx.set_wall_hits(idx, dirs, kinds, vals)
    ctx.set_collision(pb["nu"], pb["dt"], force=pb["F"], force_type=_capi.SHIFTING_VELOCITY)
    ctx.upload_populations(0, pb["f0"])
    f = pb["f0"].copy()
    ctx.step(200)
    ctx.synchronize()
    for _ in range(200):
        f = oracle_lib.stream(pb["blocks"], f)
        oracle_lib.apply_wall_hits(ost, f, None, idx, dirs, kinds, vals)
        _, ru, rc = oracle_lib.collide_advanced(ost, f, pb["nu"], pb["dt"], force=pb["F"], force_type="SHIFTING_VELOCITY")
        assert rc == 0
    assert rel_err(ctx.download_populations(0), f) <= 1e-10
    _, u = ctx.download_moments()
    assert np.max(np.abs(u - ru)) <= 1e-10 * np.max(np.abs(ru))
    ctx.step(5000)
    ctx.synchronize()
    _, u = ctx.download_moments()
    mean = common.integral_mean(pb["dofs"], pb["mesh"], 2, u[0])
    assert 0.9 * pb["u_bulk"] < mean < 1.1 * pb["u_bulk"], (mean, pb["u_bulk"])
    ctx.close()


def test_moving_walls_on_device(oracle_lib):
    """WallTest fixture (test/boundaries/WallFixture.h:44-88) on the device: non-zero VelocityNeqBounceBack terms.  Same
    trajectory as the oracle after 100 steps and the reference's bounds max |u - 0.01| < 1e-3, max |v| < 1e-5."""
    from natrium_b200 import Context, Stencil
    pb = common.moving_walls_problem(oracle_lib)
    ost, n = pb["st"], pb["dofs"].N
    st = Stencil("D2Q9", 1.0)
    idx, dirs, kinds, vals = pb["hits"]
    ctx = Context(0)
    ctx.set_stencil(st.getDirections(), st.getWeights(), st.getScaling(), st.getSpeedOfSoundSquare())
    ctx.set_layout(n, 0, False)
    _upload_oracle_blocks(ctx, pb["blocks"])
    ctx.set_wall_hits(idx, dirs, kinds, vals)
    ctx.set_collision(pb["nu"], pb["dt"])
    ctx.upload_populations(0, pb["f0"])
    ctx.step(100)
    ctx.synchronize()
    f = pb["f0"].copy()
    for _ in range(100):
        f = oracle_lib.stream(pb["blocks"], f)
        oracle_lib.apply_wall_hits(ost, f, None, idx, dirs, kinds, vals)
        oracle_lib.collide_bgk(ost, f, pb["nu"], pb["dt"])
    assert rel_err(ctx.download_populations(0), f) <= 1e-10
    rho, u = ctx.download_moments()
    assert np.max(np.abs(u[0] - 0.01)) < 1e-3 and np.max(np.abs(u[1])) < 1e-5
    assert abs(np.abs(rho).sum() / n - 1.0) < 1e-10
    ctx.close()


# ---------------------------------------------------------------------------------------------
# grid (TMA box) kernels, nb200_set_dof_grid
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["c1_tgv2d_d2q9", "tgv2d_small", "tgv3d_d3q19_small", "tgv3d_d3q19_p2", "tgv3d_d3q15", "tgv2d_d2q25", "tgv3d_d3q45"])
@pytest.mark.parametrize("numbering", ["lex", "cell"])
def test_grid_kernels_selected_and_rows_come_from_boxes(case, numbering):
    """With the grid hint the TMA box kernels drive the step; on the periodic meshes of the path every row is a box row
    (none falls back to its dictionary list), whatever the host numbering."""
    from natrium_b200 import Context, harness
    c, st, pb, dt = common.product_problem(case)
    ctx = Context(0)
    ctx.set_stencil(st.getDirections(), st.getWeights(), st.getScaling(), st.getSpeedOfSoundSquare())
    part = harness.SlabPartition(pb, st, dt)
    ctx.set_layout(part.n_owned, part.n_ghost, bool(c.get("with_g")))
    num = harness.CellNumbering(part) if numbering == "cell" else None
    ctx.set_dof_grid(*(num or part).grid_coords(), fe_order=pb.p)
    nnz = harness.upload_streaming_matrix(ctx, pb, part, st, dt, num)
    info = ctx.grid_info()
    assert info["in_use"] == 1, info
    # every (row, direction) is either a box row or taken from its dictionary list; on these tiny periodic meshes the
    # corner tiles would need up to 2^dim boxes per direction, beyond the buffer: those rows fall back, the others do not
    assert info["box_rows"] + info["generic_rows"] == (st.getQ() - 1) * part.n_owned, info
    assert info["box_rows"] >= 0.5 * (st.getQ() - 1) * part.n_owned, info
    if case in ("c1_tgv2d_d2q9", "tgv2d_d2q25"):
        assert info["generic_rows"] == 0, info
    assert info["boxes"] >= info["tiles"] and info["passes"] >= info["tiles"]
    # M 1 = 1 through the boxes (SemiLagrangian_test.cpp:519-596)
    ones = np.ones((st.getQ(), part.n_owned))
    ctx.upload_populations(0, ones)
    ctx.stream(0)
    assert np.max(np.abs(ctx.download_populations(0) - 1.0)) <= 1e-13
    ctx.close()


def test_grid_hint_errors():
    from natrium_b200 import Context, Stencil, NatriumB200Error
    st = Stencil("D2Q9", 1.0)
    ctx = Context(0)
    ctx.set_stencil(st.getDirections(), st.getWeights(), 1.0, st.getSpeedOfSoundSquare())
    with pytest.raises(NatriumB200Error):
        ctx.set_dof_grid([4, 4], np.zeros((16, 2)), 0)                      # before set_layout
    ctx.set_layout(16, 0, False)
    with pytest.raises(NatriumB200Error):
        ctx.set_dof_grid([4, 4], np.zeros((16, 2)), 0)                      # two DoFs at one grid point
    with pytest.raises(NatriumB200Error):
        ctx.set_dof_grid([4, 4, 1], np.zeros((16, 3)), 0)                   # dimension differs from the stencil's
    i = np.arange(16)
    with pytest.raises(NatriumB200Error):
        ctx.set_dof_grid([4, 3], np.stack([i % 4, i // 4], axis=1), 0)      # coordinate outside the grid
    ctx.set_dof_grid([4, 4], np.stack([i % 4, i // 4], axis=1), 0)
    with pytest.raises(NatriumB200Error):
        ctx.set_dof_order(np.arange(16))                                    # the order hint comes first
    ctx.close()


@pytest.mark.parametrize("case,with_grid", [("tgv3d_d3q19_small", False), ("tgv3d_d3q19_small", True), ("tgv3d_d3q45", False), ("tgv3d_d3q45", True)])
def test_conserved_moments_1000_steps_3d(case, with_grid, oracle_lib):
    """north_star: conserved moments (mass, momentum, energy) agree with the reference-ordered CPU run to machine precision
    over 1000 steps -- D3Q19 BGK and D3Q45 f+g (quartic, Pr 0.71, Sutherland), staged and grid kernels."""
    o = common.oracle_problem(case)
    ctx, c, st, pb, dt, part = make_ctx(case, fmt=("dict", 1e-14, False, "grid") if with_grid else None)
    with_g = bool(c.get("with_g"))
    set_collision(ctx, c, dt)
    stepper = oracle_lib.ReferenceOrderStepper(o["st"], o["blocks"], o["dofs"].N, c["nu"], dt, equilibrium=1 if with_g else 0,
                                               with_g=with_g, gamma=1.4, prandtl=0.71 if with_g else None, sutherland=with_g)
    f = o["f"].copy()
    g = o["g"].copy() if with_g else None
    ctx.upload_populations(0, f)
    if with_g:
        ctx.upload_populations(1, g)
    e, cs2 = o["st"].e, o["st"].cs2

    def sums(fa, ga):
        rho = fa.sum(axis=0)
        m = e.T @ fa
        if ga is None:
            en = (0.5 * (m ** 2).sum(axis=0) / rho).sum()
        else:      # 0.5 * (sum_i |e_i|^2 f_i / cs2 + sum_i g_i): kinetic + internal energy carried by f and g (nb200_conserved)
            en = 0.5 * (((e ** 2).sum(axis=1) @ fa) / cs2 + ga.sum(axis=0)).sum()
        return np.array([rho.sum(), m[0].sum(), m[1].sum(), m[2].sum(), en])

    done = 0
    en0 = abs(sums(f, g)[4])
    for target in (1, 10, 100, 1000):
        ctx.step(target - done)
        for _ in range(target - done):
            assert stepper.step(f, g) == 0
        done = target
        got, ref = ctx.conserved(), sums(f, g)
        assert abs(got[0] - ref[0]) <= TOL_CONS * abs(ref[0]), (target, got, ref)
        # the f-only case is a viscous Taylor-Green vortex at Re = 1: its kinetic energy decays by ten orders of magnitude
        # over the run, so the energy is compared on the scale it started from
        assert abs(got[4] - ref[4]) <= 1e-11 * max(abs(ref[4]), en0), (target, got, ref)
        scale = np.abs(e).max() * ref[0]
        assert np.max(np.abs(got[1:4] - ref[1:4])) <= TOL_CONS * scale, (target, got, ref)
        assert rel_err(ctx.download_populations(0), f) <= 1e-9, target
    ctx.synchronize()
    if with_grid:
        assert ctx.grid_info()["in_use"] == 1
    ctx.close()


@pytest.mark.parametrize("case", ["c1_tgv2d_d2q9", "tgv3d_d3q19_small"])
def test_grid_step_host_and_mixed_calls(case, oracle_lib):
    """The grid copies follow every way the populations can change: uploads, stream-only, collide-only, fused steps and
    host-buffer steps in any order give what the same calls give without the hint."""
    o = common.oracle_problem(case)
    res = []
    for fmt in (None, ("dict", 1e-14, False, "grid")):
        ctx, c, st, pb, dt, part = make_ctx(case, fmt=fmt)
        set_collision(ctx, c, dt)
        ctx.upload_populations(0, o["f"])
        ctx.step(2)
        ctx.stream(0)
        ctx.collide()
        ctx.step(1)
        n, Q = part.n_owned, st.getQ()
        fin, fout = np.ascontiguousarray(ctx.download_populations(0)), np.zeros((Q, n))
        rho, u = np.zeros(n), np.zeros((st.getD(), n))
        ctx.step_host(fin.ctypes.data, fout.ctypes.data, rho.ctypes.data, u.ctypes.data, 4)
        ctx.synchronize()
        ctx.step(2)
        ctx.upload_population(0, 1, np.ascontiguousarray(fout[1]))
        ctx.step(1)
        ctx.synchronize()
        res.append((ctx.download_populations(0), fout, rho))
        ctx.close()
    for a, b in zip(res[0], res[1]):
        assert rel_err(b, a) <= 1e-12


# ---------------------------------------------------------------------------------------------------
# ExponentialFilter on the device: nb200_set_filter / nb200_apply_filter / filtered nb200_step
# ---------------------------------------------------------------------------------------------------
def _filter_tables(p, dim, alpha=36.0, s=2.0, Nc=1, by_sum=False):
    from oracle import filter as F
    to, fr = F.projection_matrices(p, dim)
    sg, damped = F.damping(p, dim, alpha, s, Nc, by_sum)
    return to, fr, sg, damped


@pytest.mark.parametrize("cell_order", ["lexicographic", "reversed", "shuffled"])
@pytest.mark.parametrize("case,fmt", [("tgv2d_small", None), ("c1_tgv2d_d2q9", ("dict", 1e-14, True)), ("tgv3d_d3q19_p2", None),
                                      ("tgv3d_d3q19_small", ("dict", 1e-14, True, "grid")), ("tgv3d_d3q15", None),
                                      ("tgv2d_d2q25", None), ("tgv3d_d3q45", None)])
def test_exponential_filter_matches_oracle(case, fmt, cell_order, oracle_lib):
    """nb200_apply_filter against the oracle's sequential cell loop (ExponentialFilter.cpp:139-199), population by
    population as CFDSolver::filter does, for any order of the host's cell loop: the level schedule must reproduce the
    order-dependent result (a cell reads the face DoFs earlier cells wrote) exactly, not just some filtered field."""
    from oracle import filter as F
    o = common.oracle_problem(case)
    ctx, c, st, pb, dt, part = make_ctx(case, with_matrix=False, fmt=fmt)
    with_g = bool(c.get("with_g"))
    to, fr, sg, damped = _filter_tables(pb.p, pb.dim, Nc=2 if pb.p > 2 else 1)
    cd = part.cell_dofs()
    if cell_order == "reversed":
        cd = cd[::-1].copy()
    elif cell_order == "shuffled":
        cd = cd[np.random.default_rng(5).permutation(len(cd))].copy()
    ctx.set_filter(cd, to, fr, sg, interval=0)
    info = ctx.filter_info()
    assert info["cells"] == len(cd) and info["dofs_per_cell"] == (pb.p + 1) ** pb.dim and 1 < info["levels"] <= len(cd)
    rng = np.random.default_rng(11)
    f = o["f"] * (1.0 + 0.05 * rng.standard_normal(o["f"].shape))
    ctx.upload_populations(0, f)
    ctx.apply_filter(0)
    got = ctx.download_populations(0)
    ref = f.copy()
    for q in range(ref.shape[0]):
        row = np.ascontiguousarray(ref[q])
        F.apply_filter(cd, to, fr, sg, damped, row)
        ref[q] = row
    scale = np.abs(ref).max(axis=1, keepdims=True)
    assert np.max(np.abs(got - ref) / scale) <= TOL_STEP
    assert np.max(np.abs(ref - f) / scale) > 1e-4               # the filter did something
    if with_g:
        g = o["g"] * (1.0 + 0.05 * rng.standard_normal(o["g"].shape))
        ctx.upload_populations(1, g)
        ctx.apply_filter(1)
        gg = ctx.download_populations(1)
        refg = g.copy()
        for q in range(refg.shape[0]):
            row = np.ascontiguousarray(refg[q])
            F.apply_filter(cd, to, fr, sg, damped, row)
            refg[q] = row
        assert np.max(np.abs(gg - refg) / np.abs(refg).max(axis=1, keepdims=True)) <= TOL_STEP
        assert rel_err(ctx.download_populations(0), got) == 0.0      # f untouched by the g pass
    ctx.close()


@pytest.mark.parametrize("case,fmt,interval", [("c1_tgv2d_d2q9", None, 1), ("tgv3d_d3q19_p2", ("dict", 1e-14, False, "grid"), 2),
                                              ("tgv3d_d3q19_small", ("dict", 1e-14, True), 3), ("tgv2d_d2q25", None, 2)])
def test_filtered_step_matches_oracle(case, fmt, interval, oracle_lib):
    """nb200_step with a filter: CFDSolver::run order stream -> filter (when m_i % interval == 0) -> collide
    (CFDSolver.cpp:884-888; compressibleFilter for f + g), 6 steps against the oracle from the same start; the steps in
    between run the fused kernel."""
    from oracle import filter as F
    o = common.oracle_problem(case)
    ctx, c, st, pb, dt, part = make_ctx(case, fmt=fmt)
    with_g = bool(c.get("with_g"))
    set_collision(ctx, c, dt)
    to, fr, sg, damped = _filter_tables(pb.p, pb.dim, alpha=5.0, s=4.0, Nc=pb.p)        # only the highest mode, mildly
    cd = part.cell_dofs()
    ctx.set_filter(cd, to, fr, sg, interval=interval)
    f = o["f"].copy()
    g = o["g"].copy() if with_g else None
    ctx.upload_populations(0, f)
    if with_g:
        ctx.upload_populations(1, g)
    steps = 6
    ctx.step(steps)
    ctx.synchronize()
    assert ctx.filter_info()["iteration"] == steps

    def filt(a):
        for q in range(a.shape[0]):
            row = np.ascontiguousarray(a[q])
            F.apply_filter(cd, to, fr, sg, damped, row)
            a[q] = row
    unfiltered = f.copy()
    for i in range(1, steps + 1):
        f = oracle_lib.stream(o["blocks"], f)
        if with_g:
            g = oracle_lib.stream(o["blocks"], g)
        if i % interval == 0:
            filt(f)
            if with_g:
                filt(g)
        if with_g:
            assert oracle_lib.collide_bgk_fg(o["st"], f, g, c["nu"], dt, equilibrium=1, gamma=1.4, prandtl=0.71, sutherland=True)[-1] == 0
        else:
            assert oracle_lib.collide_bgk(o["st"], f, c["nu"], dt)[-1] == 0
    got = ctx.download_populations(0)
    assert rel_err(got, f) <= 10 * TOL_STEP, rel_err(got, f)
    if with_g:
        assert rel_err(ctx.download_populations(1), g) <= 10 * TOL_STEP
    # and it is not the unfiltered run
    ctx.set_filter(None, None, None, None)
    ctx.upload_populations(0, unfiltered)
    if with_g:
        ctx.upload_populations(1, o["g"])
    ctx.step(steps)
    assert rel_err(ctx.download_populations(0), f) > 1e-9
    ctx.close()


def test_filter_errors():
    from natrium_b200 import NatriumB200Error
    ctx, c, st, pb, dt, part = make_ctx("tgv2d_small", with_matrix=False)
    to, fr, sg, _ = _filter_tables(pb.p, pb.dim)
    cd = part.cell_dofs()
    with pytest.raises(NatriumB200Error):
        ctx.apply_filter(0)                                    # nothing set
    bad = cd.copy(); bad[0, 0] = part.n_owned + 5
    with pytest.raises(NatriumB200Error):
        ctx.set_filter(bad, to, fr, sg)
    bad = cd.copy(); bad[1, 1] = bad[1, 0]
    with pytest.raises(NatriumB200Error):
        ctx.set_filter(bad, to, fr, sg)                        # a DoF twice in one cell
    ctx.set_filter(cd, to, fr, sg)
    with pytest.raises(NatriumB200Error):
        ctx.apply_filter(1)                                    # no g in the layout
    ctx.set_filter(None, None, None, None)
    assert ctx.filter_info()["cells"] == 0
    ctx.close()


def test_host_mirror_solver_with_filter(oracle_lib):
    """The reference-shaped host API with a filter (CFDSolver::filter between stream and collide, CFDSolver.cpp:884-888):
    natrium_b200.ExponentialFilter builds the tables the reference builds with deal.II, CFDSolver.setFilter hands them over,
    run() filters inside nb200_step; compared with the oracle's stream -> filter -> collide loop that uses the ORACLE's own
    tables (oracle/filter.py), so the host mirror's tables are checked end to end as well."""
    from oracle import filter as F
    from natrium_b200 import CFDSolver, ExponentialFilter, SolverConfiguration, harness
    case = "tgv2d_small"
    o = common.oracle_problem(case)
    c = o["c"]
    cfg = SolverConfiguration()
    cfg.setStencil("D2Q9")
    cfg.setStencilScaling(c["scaling"])
    cfg.setSedgOrderOfFiniteElement(c["p"])
    cfg.setCFL(c["cfl"])
    pb = harness.CartesianProblem(c["dim"], c["cells"], c["p"])
    solver = CFDSolver(cfg, pb, c["nu"])
    solver.setInitialFields(o["rho"], o["u"])
    alpha, s, Nc = 8.0, 4.0, c["p"]
    solver.setFilter(ExponentialFilter(alpha, s, Nc, False, c["p"], c["dim"]), interval=2)
    solver.run(4)
    to, fr = F.projection_matrices(c["p"], c["dim"])
    sg, damped = F.damping(c["p"], c["dim"], alpha, s, Nc)
    cd = solver.getAdvectionOperator().getPartition().cell_dofs()
    f = o["f"].copy()
    oracle_lib.collide_bgk(o["st"], f, c["nu"], o["dt"])          # run(): collide once before the loop
    for i in range(1, 5):
        f = oracle_lib.stream(o["blocks"], f)
        if i % 2 == 0:
            for q in range(f.shape[0]):
                row = np.ascontiguousarray(f[q])
                F.apply_filter(cd, to, fr, sg, damped, row)
                f[q] = row
        assert oracle_lib.collide_bgk(o["st"], f, c["nu"], o["dt"])[-1] == 0
    assert rel_err(solver.getF().to_host(), f) <= 1e-11
    solver.ctx.close()
