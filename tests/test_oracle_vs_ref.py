"""The oracle restatement against the REFERENCE's own code (oracle/_ref/libnatrium_ref.so, compiled from
/root/reference by oracle/Makefile.ref with the stand-in headers of oracle/ref_stubs).

Every scheme / stencil / flag combination the GPU parity tests use (tests/test_gpu_parity.py) is run through
natrium::selectCollision<dim> (L/collision_advanced/CollisionSelection.h:69-272) or the legacy
CollisionModel::collideAll (L/collision/KBCStandard.cpp, MRTEntropic.cpp) and through the oracle on identical inputs.
Tolerance: 1e-15 relative to the largest population (in practice the results are bit-identical: same operation order,
same -ffp-contract=off); the GPU tests then compare the CUDA path with the oracle at 1e-12.

CPU only (no `gpu` marker).  The library travels to the GPU box prebuilt; here it is rebuilt from the sources.
"""
import numpy as np
import pytest

from oracle import cpu, ref

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built and /root/reference absent")

TOL = 1e-15


def synthetic_populations(Q, n):
    """The reference's own test population (BGKStandard_test.cpp:358-508): 1.5 + sin(1.5 i) + 0.001 + (0.5 cos j)^2."""
    i = np.arange(Q, dtype=np.float64)[:, None]
    j = np.arange(n, dtype=np.float64)[None, :]
    return 1.5 + np.sin(1.5 * i) + 0.001 + (0.5 * np.cos(j)) ** 2 + 0 * i


def close(a, b, tol=TOL):
    scale = max(1e-300, float(np.max(np.abs(b))))
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b)))) <= tol * scale


@pytest.mark.parametrize("name", ["D2Q9", "D3Q19", "D3Q15", "D2Q25H", "D3Q45"])
@pytest.mark.parametrize("scaling", [1.0, 2.5, np.sqrt(3) / 0.05])
def test_stencil_tables(name, scaling):
    """a11: oracle/stencils.py and the product's natrium_b200/stencils.py against L/stencils/*.cpp compiled here."""
    from oracle import stencils as ost
    from natrium_b200.stencils import Stencil
    e, w, cs2, mx, opp = ref.stencil(name, scaling)
    oe, ow, ocs2, omx = ost.make(name, scaling)
    assert np.array_equal(e, oe) and np.array_equal(w, ow) and cs2 == ocs2 and mx == omx
    ps = Stencil(name, scaling)
    assert np.array_equal(ps.getDirections(), e) and np.array_equal(ps.getWeights(), w)
    assert ps.getSpeedOfSoundSquare() == cs2 and ps.getMaxParticleVelocityMagnitude() == mx
    for i in range(len(w)):
        assert np.array_equal(e[opp[i]], -e[i])


@pytest.mark.parametrize("in_init", [False, True])
@pytest.mark.parametrize("stencil,eq", [("D2Q9", 0), ("D2Q9", 1), ("D3Q19", 0), ("D3Q15", 0), ("D2Q25H", 1), ("D3Q45", 0), ("D3Q45", 1)])
def test_bgk_f(stencil, eq, in_init):
    """a7/a9: BGKCollision::relax rows of selectCollision(f), incl. the f-only D3Q45 'quartic' row that instantiates
    BGKEquilibrium (CollisionSelection.h:199)."""
    scaling = 1.0 if stencil in ("D2Q25H", "D3Q45") else 2.5
    st = cpu.Stencil(stencil, scaling)
    n, dt = 1000, 0.1
    nu = 0.9 * dt * st.cs2
    f = synthetic_populations(st.Q, n) * st.w[:, None]
    u0 = 0.05 * scaling * np.vstack([np.sin(np.arange(n) + d) for d in range(st.D)])
    a, b = f.copy(), f.copy()
    r = ref.select_collision(stencil, scaling, a, nu, dt, equilibrium=["BGK_EQUILIBRIUM", "QUARTIC_EQUILIBRIUM"][eq],
                             in_init=in_init, u_init=u0 if in_init else None)
    rho, u, rc = cpu.collide_bgk(st, b, nu, dt, equilibrium=0 if stencil == "D3Q45" else eq, in_init=in_init,
                                 u_init=u0.copy() if in_init else None)
    assert r["status"] == 0 and rc == 0, r["message"]
    assert close(b, a) and close(rho, r["rho"]) and close(u, r["u"])


@pytest.mark.parametrize("stencil,eq,prandtl,sutherland", [("D2Q25H", 1, 0.71, True), ("D2Q25H", 1, None, False),
                                                           ("D2Q25H", 0, None, False), ("D3Q45", 1, 0.71, True),
                                                           ("D3Q45", 1, None, False), ("D3Q45", 1, 0.7, False)])
def test_bgk_fg(stencil, eq, prandtl, sutherland):
    """a8: relaxWithG (quartic equilibrium with temperature, Prandtl correction, Sutherland law, shock sensor)."""
    from natrium_b200 import harness
    from natrium_b200.stencils import Stencil
    st = cpu.Stencil(stencil, 1.0)
    n, dt, nu, gamma = 777, 0.05, 0.002, 1.4
    rng = np.random.default_rng(7)
    rho = 1.0 + 0.1 * rng.standard_normal(n)
    u = 0.1 * rng.standard_normal((st.D, n))
    T = 1.0 + 0.05 * rng.standard_normal(n)
    f, g = harness.quartic_equilibrium_distributions(Stencil(stencil, 1.0), rho, u, T, gamma)
    f *= 1.0 + 0.01 * rng.standard_normal(f.shape)
    g *= 1.0 + 0.01 * rng.standard_normal(g.shape)
    af, ag, bf, bg = f.copy(), g.copy(), f.copy(), g.copy()
    r = ref.select_collision(stencil, 1.0, af, nu, dt, equilibrium=["BGK_EQUILIBRIUM", "QUARTIC_EQUILIBRIUM"][eq], g=ag,
                             gamma=gamma, prandtl=prandtl, sutherland=sutherland)
    orho, ou, oT, os_, rc = cpu.collide_bgk_fg(st, bf, bg, nu, dt, equilibrium=eq, gamma=gamma, prandtl=prandtl, sutherland=sutherland)
    assert r["status"] == 0 and rc == 0, r["message"]
    assert close(bf, af) and close(bg, ag)
    assert close(orho, r["rho"]) and close(ou, r["u"]) and close(oT, r["T"]) and close(os_, r["sensor"])


ADVANCED = [("D2Q9", "BGK_REGULARIZED", "DELLAR_D2Q9", "RELAX_FULL"), ("D3Q15", "BGK_REGULARIZED", "DELLAR_D2Q9", "RELAX_FULL"),
            ("D3Q19", "BGK_REGULARIZED", "DELLAR_D2Q9", "RELAX_FULL"), ("D2Q9", "MRT_STANDARD", "DELLAR_D2Q9", "RELAX_FULL"),
            ("D2Q9", "MRT_STANDARD", "DELLAR_D2Q9", "DELLAR_RELAX_ONLY_N"), ("D2Q9", "MRT_STANDARD", "LALLEMAND_D2Q9", "RELAX_FULL"),
            ("D3Q19", "MRT_STANDARD", "DHUMIERES_D3Q19", "RELAX_FULL"), ("D3Q19", "MRT_STANDARD", "DHUMIERES_D3Q19", "RELAX_DHUMIERES_PAPER")]


@pytest.mark.parametrize("in_init", [False, True])
@pytest.mark.parametrize("stencil,scheme,basis,relax", ADVANCED)
def test_regularized_and_mrt(stencil, scheme, basis, relax, in_init):
    """f2: Regularized::relax / MultipleRelaxationTime::relax with the tables of AuxiliaryMRTFunctions.cpp."""
    scaling = 2.5
    st = cpu.Stencil(stencil, scaling)
    n, dt = 1000, 0.1
    nu = 0.9 * dt * st.cs2
    f = synthetic_populations(st.Q, n) * st.w[:, None]
    u0 = 0.05 * scaling * np.vstack([np.sin(np.arange(n) + d) for d in range(st.D)])
    a, b = f.copy(), f.copy()
    r = ref.select_collision(stencil, scaling, a, nu, dt, scheme=scheme, in_init=in_init, u_init=u0 if in_init else None,
                             mrt_basis=basis, relax_mode=relax)
    rho, u, rc = cpu.collide_advanced(st, b, nu, dt, scheme=scheme, in_init=in_init, u_init=u0.copy() if in_init else None,
                                      mrt_basis=basis, relax_mode=relax)
    assert r["status"] == 0 and rc == 0, r["message"]
    assert close(b, a) and close(rho, r["rho"]) and close(u, r["u"])


@pytest.mark.parametrize("basis,Q,relax", [("DELLAR_D2Q9", 9, "RELAX_FULL"), ("DELLAR_D2Q9", 9, "DELLAR_RELAX_ONLY_N"),
                                           ("LALLEMAND_D2Q9", 9, "RELAX_FULL"), ("DHUMIERES_D3Q19", 19, "RELAX_FULL"),
                                           ("DHUMIERES_D3Q19", 19, "RELAX_DHUMIERES_PAPER")])
def test_mrt_tables(basis, Q, relax):
    """make_M / make_T / make_diag: golden fixture (oracle), product host mirror (natrium_b200/mrt.py) and reference agree."""
    from natrium_b200 import mrt
    tau = 0.83
    M, T, om = ref.mrt_tables(Q, basis, relax, tau)
    oM, oT = cpu.mrt_tables(basis)
    assert np.array_equal(M, oM) and np.array_equal(T, oT)
    assert np.array_equal(om, cpu.mrt_diag(tau, basis, relax))
    b, r = getattr(mrt, basis), getattr(mrt, relax)
    assert np.allclose(mrt.make_M(b), M, rtol=0, atol=1e-15) and np.allclose(mrt.make_T(b), T, rtol=0, atol=1e-15)
    assert np.array_equal(mrt.make_diag(tau, b, r), om)


@pytest.mark.parametrize("in_init", [False, True])
@pytest.mark.parametrize("force_type", ["SHIFTING_VELOCITY", "EXACT_DIFFERENCE"])
@pytest.mark.parametrize("stencil,scheme,basis", [("D2Q9", "BGK_STANDARD", "DELLAR_D2Q9"), ("D3Q19", "BGK_STANDARD", "DELLAR_D2Q9"),
                                                  ("D3Q19", "BGK_REGULARIZED", "DELLAR_D2Q9"), ("D2Q9", "MRT_STANDARD", "LALLEMAND_D2Q9")])
def test_forced_f(stencil, scheme, basis, force_type, in_init):
    """f2: applyMacroscopicForces / applyForces / postCollisionApplyForces (AuxiliaryCollisionFunctions.h:332-417)."""
    scaling = 2.0
    st = cpu.Stencil(stencil, scaling)
    n, dt = 640, 0.1
    nu = 0.9 * dt * st.cs2
    F = np.array([1e-2, -2e-2, 5e-3])[:st.D]
    f = synthetic_populations(st.Q, n) * st.w[:, None]
    u0 = 0.05 * scaling * np.vstack([np.sin(np.arange(n) + d) for d in range(st.D)])
    a, b = f.copy(), f.copy()
    r = ref.select_collision(stencil, scaling, a, nu, dt, scheme=scheme, in_init=in_init, u_init=u0 if in_init else None,
                             force=F, force_type=force_type, mrt_basis=basis)
    rho, u, rc = cpu.collide_advanced(st, b, nu, dt, scheme=scheme, in_init=in_init, u_init=u0.copy() if in_init else None,
                                      force=F, force_type=force_type, mrt_basis=basis)
    assert r["status"] == 0 and rc == 0, r["message"]
    assert close(b, a) and close(rho, r["rho"]) and close(u, r["u"])


@pytest.mark.parametrize("force_type", ["SHIFTING_VELOCITY", "EXACT_DIFFERENCE"])
@pytest.mark.parametrize("stencil", ["D2Q25H", "D3Q45"])
def test_forced_fg(stencil, force_type):
    """The channel configuration's collision: f + g, Pr 0.7, Sutherland, EXACT_DIFFERENCE (step-turbulent-channel.cpp:159-177)."""
    from natrium_b200 import harness
    from natrium_b200.stencils import Stencil
    st = cpu.Stencil(stencil, 1.0)
    n, dt, nu, gamma = 500, 0.05, 0.002, 1.4
    rng = np.random.default_rng(11)
    rho = 1.0 + 0.1 * rng.standard_normal(n)
    u = 0.1 * rng.standard_normal((st.D, n))
    T = 1.0 + 0.05 * rng.standard_normal(n)
    F = np.array([3e-2, 0.0, -1e-2])[:st.D]
    f, g = harness.quartic_equilibrium_distributions(Stencil(stencil, 1.0), rho, u, T, gamma)
    f *= 1.0 + 0.01 * rng.standard_normal(f.shape)
    g *= 1.0 + 0.01 * rng.standard_normal(g.shape)
    af, ag, bf, bg = f.copy(), g.copy(), f.copy(), g.copy()
    r = ref.select_collision(stencil, 1.0, af, nu, dt, equilibrium="QUARTIC_EQUILIBRIUM", g=ag, gamma=gamma, prandtl=0.7,
                             sutherland=True, force=F, force_type=force_type)
    orho, ou, oT, os_, rc = cpu.collide_bgk_fg_forced(st, bf, bg, nu, dt, F, force_type, gamma=gamma, prandtl=0.7, sutherland=True)
    assert r["status"] == 0 and rc == 0, r["message"]
    assert close(bf, af) and close(bg, ag)
    assert close(orho, r["rho"]) and close(ou, r["u"]) and close(oT, r["T"]) and close(os_, r["sensor"])


@pytest.mark.parametrize("in_init", [False, True])
@pytest.mark.parametrize("stencil,scheme", [("D2Q9", "KBC_STANDARD"), ("D3Q15", "KBC_STANDARD"), ("D3Q19", "MRT_ENTROPIC")])
def test_legacy_entropic(stencil, scheme, in_init):
    """a10: KBCStandard::collideAllD2Q9 / D3Q15 (KBCStandard.cpp:88-1028), MRTEntropic::collideAllD3Q19 (MRTEntropic.cpp:167-305)."""
    scaling = 2.5
    st = cpu.Stencil(stencil, scaling)
    n, dt = 1000, 0.1
    nu = 0.9 * dt * st.cs2
    f = synthetic_populations(st.Q, n) * st.w[:, None]
    u0 = 0.05 * scaling * np.vstack([np.sin(np.arange(n) + d) for d in range(st.D)])
    a, b = f.copy(), f.copy()
    r = ref.legacy_collide(stencil, scaling, scheme, a, nu, dt, in_init=in_init, u_init=u0 if in_init else None)
    rho, u, rc = cpu.collide_entropic(st, b, nu, dt, scheme, in_init=in_init, u_init=u0.copy() if in_init else None)
    assert r["status"] == 0 and rc == 0, r["message"]
    assert close(b, a) and close(rho, r["rho"]) and close(u, r["u"])


@pytest.mark.parametrize("stencil", ["D2Q9", "D3Q19"])
def test_legacy_bgk_equals_advanced_bgk(stencil):
    """SURVEY 8c's free cross-check, now on the reference's own two implementations: legacy BGKStandard::collideAll
    (scaled velocities, prefactor -1/(tau+1/2)) and collision_advanced BGK agree to round-off; the oracle restates the latter."""
    scaling = 2.5
    st = cpu.Stencil(stencil, scaling)
    n, dt = 500, 0.1
    nu = 0.9 * dt * st.cs2
    f = synthetic_populations(st.Q, n) * st.w[:, None]
    a, b = f.copy(), f.copy()
    r1 = ref.legacy_collide(stencil, scaling, "BGK_STANDARD", a, nu, dt)
    r2 = ref.select_collision(stencil, scaling, b, nu, dt)
    assert r1["status"] == 0 and r2["status"] == 0
    assert close(a, b, 1e-13) and close(r1["rho"], r2["rho"], 1e-14)


def test_legacy_equilibrium_used_by_initialisation():
    """a12: f = f_eq(rho0, u0) through BGKStandard::getEquilibriumDistribution (BGKStandard.cpp:23-41) vs oracle and harness."""
    from natrium_b200 import harness
    from natrium_b200.stencils import Stencil
    for name, scaling in (("D2Q9", np.sqrt(3) / 0.05), ("D3Q19", np.sqrt(3) / 0.05), ("D3Q15", 2.0)):
        st = cpu.Stencil(name, scaling)
        rng = np.random.default_rng(3)
        for _ in range(5):
            rho = 1.0 + 0.1 * rng.standard_normal()
            u = 0.3 * rng.standard_normal(st.D)
            want = ref.legacy_feq(name, scaling, rho, u)
            assert close(cpu.legacy_feq(st, rho, u), want)
            got = harness.equilibrium_distributions(Stencil(name, scaling), np.array([rho]), u[:, None])[:, 0]
            assert close(got, want, 1e-14)


def test_error_paths_match():
    """Density below 1e-10 -> CollisionException; an unlisted (stencil, scheme, equilibrium) row -> 'not implemented yet';
    a forced problem with NO_FORCING -> NATriuMException (CollisionSelection.h:102-110, Aux...h:53-56,334-338)."""
    st = cpu.Stencil("D2Q9", 1.0)
    f = np.zeros((9, 4))
    r = ref.select_collision("D2Q9", 1.0, f.copy(), 0.1, 0.1)
    assert r["status"] == -1 and "Densities too small" in r["message"]
    _, _, rc = cpu.collide_bgk(st, f.copy(), 0.1, 0.1)
    assert rc != 0
    f = np.ones((19, 4)) / 19
    r = ref.select_collision("D3Q19", 1.0, f.copy(), 0.1, 0.1, equilibrium="QUARTIC_EQUILIBRIUM")
    assert r["status"] == -1 and "not implemented yet" in r["message"]
    r = ref.select_collision("D3Q19", 1.0, f.copy(), 0.1, 0.1, force=[1e-3, 0, 0], force_type="NO_FORCING")
    assert r["status"] == -2 and "forcing was switched off" in r["message"]
    r = ref.legacy_collide("D3Q19", 1.0, "KBC_STANDARD", f.copy(), 0.1, 0.1)
    assert r["status"] == -1 and "only implemented for D2Q9 and D3Q15" in r["message"]


def test_thermal_bounce_back_point():
    """f1: the oracle's wall-hit replay (kind 1 = ThermalBounceBack, T_w = 0.85) against the reference's own arithmetic."""
    from natrium_b200 import harness
    from natrium_b200.stencils import Stencil
    st = cpu.Stencil("D3Q45", 1.0)
    n, gamma = 40, 1.4
    rng = np.random.default_rng(5)
    rho = 1.0 + 0.1 * rng.standard_normal(n)
    u = 0.1 * rng.standard_normal((3, n))
    T = np.where(np.arange(n) % 4 == 0, 0.85, 1.0 + 0.05 * rng.standard_normal(n))     # every 4th DoF already at T_w
    f, g = harness.quartic_equilibrium_distributions(Stencil("D3Q45", 1.0), rho, u, T, gamma)
    f[:, 1::4] *= 1.0 + 0.01 * rng.standard_normal(f[:, 1::4].shape)
    af, ag = f.copy(), g.copy()
    idx = np.arange(n, dtype=np.int32)
    cpu.apply_wall_hits(st, af, ag, idx, np.ones(n, dtype=np.int32), np.ones(n, dtype=np.int32), np.full(n, 0.85))
    changed = 0
    for i in range(n):
        fi, gi = np.ascontiguousarray(f[:, i]), np.ascontiguousarray(g[:, i])
        changed += ref.thermal_wall_point(1.0, 0.85, fi, gi)
        assert close(af[:, i], fi) and close(ag[:, i], gi)
    assert 0 < changed < n


# ---------------------------------------------------------------------------------------------------
# f4: the reference's own ExponentialFilter (L/smoothing/ExponentialFilter.cpp compiled against oracle/ref_stubs_filter)
# ---------------------------------------------------------------------------------------------------
filter_ref = pytest.mark.skipif(not ref.filter_available(), reason="oracle/_ref filter library not built and /root/reference absent")


@filter_ref
@pytest.mark.parametrize("p,dim", [(4, 1), (2, 2), (4, 2), (2, 3), (3, 3), (4, 3)])
def test_filter_projection_matrices_vs_reference(p, dim):
    """makeProjectionMatrices (ExponentialFilter.cpp:30-68) as compiled from the reference against oracle/filter.py's
    restatement and the product's host mirror (natrium_b200.host.ExponentialFilter).  The three invert different matrices with
    different algorithms (Gauss-Jordan on the quadrature sums / LAPACK on the same / LAPACK on the nodal Vandermonde), so the
    bar is 1e-11 relative to the largest entry; to @ from = 1 to 1e-12."""
    from oracle import filter as F
    from natrium_b200 import host
    to_r, fr_r = ref.exponential_filter(dim, p, 36.0, 2.0, 1)
    to_o, fr_o = F.projection_matrices(p, dim)
    h = host.ExponentialFilter(36.0, 2.0, 1, False, p, dim)
    n = (p + 1) ** dim
    assert np.abs(to_r @ fr_r - np.eye(n)).max() <= 1e-12
    for a, b in ((to_o, to_r), (fr_o, fr_r), (h.getProjectToLegendre(), to_r), (h.getProjectFromLegendre(), fr_r)):
        assert np.abs(a - b).max() <= 1e-11 * np.abs(b).max()


@filter_ref
@pytest.mark.parametrize("p,dim,alpha,s,Nc,by_sum", [(4, 1, 36.0, 2.0, 1, False), (3, 2, 36.0, 2.0, 1, False), (4, 2, 10.0, 4.0, 2, True),
                                                    (2, 3, 36.0, 2.0, 1, False), (3, 3, 8.0, 4.0, 2, False), (4, 3, 36.0, 2.0, 1, False),
                                                    (3, 3, 10.0, 2.0, 2, True)])
def test_filter_damping_per_mode_vs_reference(p, dim, alpha, s, Nc, by_sum):
    """makeDegreeVectors + the damping formula inside applyFilter (ExponentialFilter.cpp:96-137,171-183): one cell whose nodal
    values are Legendre mode i comes back scaled by sigma_i, which reads the factor the REFERENCE applies to every mode off
    its own compiled code -- including the 3-d degree vector's operator-precedence quirk (:123, iy evaluates to iz) that
    oracle/filter.py and the host mirror restate on purpose."""
    from oracle import filter as F
    from natrium_b200 import host
    n = (p + 1) ** dim
    to_r, fr_r = ref.exponential_filter(dim, p, alpha, s, Nc, by_sum)
    sigma_ref = np.zeros(n)
    cd = np.arange(n, dtype=np.int32)[None, :]
    for i in range(n):
        v = np.ascontiguousarray(fr_r[:, i])
        ref.exponential_filter(dim, p, alpha, s, Nc, by_sum, cell_dofs=cd, v=v)
        k = int(np.argmax(np.abs(fr_r[:, i])))
        sigma_ref[i] = v[k] / fr_r[k, i]
        assert np.abs(v - sigma_ref[i] * fr_r[:, i]).max() <= 1e-11 * np.abs(fr_r[:, i]).max()
    sg, damped = F.damping(p, dim, alpha, s, Nc, by_sum)
    assert np.abs(sg - sigma_ref).max() <= 1e-11
    assert np.abs(host.ExponentialFilter(alpha, s, Nc, by_sum, p, dim).sigma - sigma_ref).max() <= 1e-11
    if dim == 3:
        sg_intended, _ = F.damping(p, dim, alpha, s, Nc, by_sum, reference_quirk=False)
        assert np.abs(sg_intended - sigma_ref).max() > 1e-3          # the reference really has the quirk


@filter_ref
@pytest.mark.parametrize("dim,cells,p,order", [(2, [4, 3], 3, "lex"), (2, [5, 4], 2, "reversed"), (3, [3, 2, 2], 2, "shuffled"), (3, [2, 2, 3], 4, "lex"),
                                               (3, [3, 3, 2], 3, "shuffled")])
def test_filter_cell_loop_vs_reference(dim, cells, p, order):
    """applyFilter (ExponentialFilter.cpp:139-199) on a mesh of cells that share their face DoFs: the oracle's C loop
    (orc_exponential_filter, fed with the REFERENCE's matrices and the oracle's damping) against the reference's own loop, for
    lexicographic, reversed and shuffled cell orders (the result depends on the order).  <= 1e-14 relative."""
    from oracle import filter as F
    from natrium_b200 import harness
    from natrium_b200.stencils import Stencil
    st = Stencil("D2Q9" if dim == 2 else "D3Q19", 3.0)
    pb = harness.CartesianProblem(dim, cells, p)
    part = harness.SlabPartition(pb, st, pb.timestep(st, 0.4))
    cd = part.cell_dofs()
    if order == "reversed":
        cd = cd[::-1].copy()
    elif order == "shuffled":
        cd = cd[np.random.default_rng(3).permutation(len(cd))].copy()
    alpha, s, Nc = 6.0, 2.0, 1
    to_r, fr_r = ref.exponential_filter(dim, p, alpha, s, Nc)
    sg, damped = F.damping(p, dim, alpha, s, Nc)
    v0 = np.random.default_rng(11).standard_normal(pb.N)
    v_ref = v0.copy()
    ref.exponential_filter(dim, p, alpha, s, Nc, cell_dofs=cd, v=v_ref)
    v_orc = F.apply_filter(cd, to_r, fr_r, sg, damped, v0.copy())
    assert np.abs(v_orc - v_ref).max() <= 1e-14 * np.abs(v_ref).max()
    assert np.abs(v_ref - v0).max() > 0.1
    # and with the oracle's own matrices: same field up to the conditioning of the projections
    to_o, fr_o = F.projection_matrices(p, dim)
    v_own = F.apply_filter(cd, to_o, fr_o, sg, damped, v0.copy())
    assert np.abs(v_own - v_ref).max() <= 1e-11 * np.abs(v_ref).max()


# ---------------------------------------------------------------------------------------------------
# f1: the reference's own ThermalBounceBack (L/boundaries/ThermalBounceBack.cpp compiled against oracle/ref_stubs_walls)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.skipif(not ref.walls_available(), reason="oracle/_ref walls library not built and /root/reference absent")
@pytest.mark.parametrize("scaling,wall_T", [(1.0, 0.85), (1.0, 1.0), (2.5, 0.85)])
def test_thermal_bounce_back_vs_reference(scaling, wall_T):
    """orc_apply_wall_hits (thermal kind) against ThermalBounceBack<3>::calculateBoundaryValues as compiled from the reference:
    D3Q45 f and g near a quartic equilibrium with a temperature field that differs from the wall temperature at some DoFs and
    equals it (within the reference's 1e-5 switch) at others; a hit list with repeated destinations (several directions hit the
    same wall DoF: the second application starts from the first one's output) replayed in order.  <= 1e-14 relative."""
    from oracle import fields
    st = cpu.Stencil("D3Q45", scaling)
    n = 64
    rng = np.random.default_rng(4)
    rho = 1.0 + 0.05 * rng.standard_normal(n)
    u = 0.03 * scaling * rng.standard_normal((3, n))
    T = wall_T + np.where(np.arange(n) % 3 == 0, 0.0, 0.05 * rng.standard_normal(n))
    f, g = fields.quartic_equilibrium_init(st.e, st.w, st.cs2, st.scaling, rho, u, T, 1.4)
    f = np.ascontiguousarray(f * (1.0 + 1e-3 * rng.standard_normal(f.shape)))
    g = np.ascontiguousarray(g * (1.0 + 1e-3 * rng.standard_normal(g.shape)))
    idx = np.concatenate([rng.integers(0, n, 40), np.arange(0, n, 5), np.arange(0, n, 5)]).astype(np.int32)
    dirs = rng.integers(1, 45, len(idx)).astype(np.int32)
    f_ref, g_ref = f.copy(), g.copy()
    ref.thermal_bounce_back(scaling, f_ref, g_ref, idx, dirs, wall_T)
    f_orc, g_orc = f.copy(), g.copy()
    assert cpu.apply_wall_hits(st, f_orc, g_orc, idx, dirs, np.ones(len(idx), dtype=np.int32), np.full(len(idx), wall_T)) == 0
    assert np.abs(f_ref - f).max() > 1e-6 and np.abs(g_ref - g).max() > 1e-6          # the wall did something
    assert np.abs(f_orc - f_ref).max() <= 1e-14 * np.abs(f_ref).max()
    assert np.abs(g_orc - g_ref).max() <= 1e-14 * np.abs(g_ref).max()
    untouched = np.setdiff1d(np.arange(n), idx)
    assert np.array_equal(f_ref[:, untouched], f[:, untouched]) and np.array_equal(g_ref[:, untouched], g[:, untouched])
