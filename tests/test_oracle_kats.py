"""Pins the CPU oracle against the reference's own known-answer / self-consistency tests
(SURVEY.md section 4).  The reference ships no golden arrays, so these are the properties and
cross-path identities its Boost.Test suite asserts, restated with the same inputs and tolerances.
CPU only."""
import math

import numpy as np
import pytest

from oracle import assembly, cpu, fields
from tests import common


def synthetic_populations(Q, n):
    """BGKStandard_test.cpp:373-380: f_i(j) = 1.5 + sin(1.5 i) + 0.001 + i/(i+1) + (0.5 cos j)^2 with integer i/(i+1) = 0."""
    i = np.arange(Q, dtype=np.float64)[:, None]
    j = np.arange(n, dtype=np.float64)[None, :]
    return 1.5 + np.sin(1.5 * i) + 0.001 + (0.5 * np.cos(j)) ** 2 + 0 * i


@pytest.mark.parametrize("name", ["D2Q9", "D3Q19", "D3Q15"])
def test_collide_all_equals_collide_single_point(name, oracle_lib):
    """BGKStandard_collideAllD2Q9/D3Q19/D3Q15_test (BGKStandard_test.cpp:358-590): the vectorised
    collideAll equals the scalar legacy collideSinglePoint to 1e-10 % (= 1e-12 relative); tau = 0.9, dt = 0.1.
    Here: collision_advanced BGK (unscaled e, 1/tau with tau = tau_legacy + 0.5) == legacy BGKStandard
    (scaled e, prefactor -1/(tau_legacy + 0.5)), the oracle-internal cross-check of SURVEY 8(c)."""
    st = oracle_lib.Stencil(name, 5.0)
    n, dt, tau_legacy = 10, 0.1, 0.9
    nu = tau_legacy * dt * st.cs2
    f = synthetic_populations(st.Q, n)
    got = f.copy()
    _, _, rc = oracle_lib.collide_bgk(st, got, nu, dt)
    assert rc == 0
    for j in range(n):
        ref = oracle_lib.legacy_collide_single_point(st, f[:, j], tau_legacy)
        assert np.max(np.abs(got[:, j] - ref) / np.abs(ref)) <= 1e-12


@pytest.mark.parametrize("name", ["D2Q9", "D3Q19", "D3Q15", "D2Q25H", "D3Q45"])
def test_collision_invariants(name, oracle_lib):
    """BGKStandardCollisionInvariants_test (:311-356): rho and rho*u conserved to 1e-15 (relative);
    f_eq is a fixed point of the collision."""
    st = oracle_lib.Stencil(name, 1.0)
    f = synthetic_populations(st.Q, 50) * st.w[:, None]
    before = f.copy()
    oracle_lib.collide_bgk(st, f, 0.03, 0.1, equilibrium=1 if name == "D2Q25H" else 0)
    assert np.max(np.abs(f.sum(0) - before.sum(0)) / before.sum(0)) <= 5e-15
    assert np.max(np.abs(st.e.T @ f - st.e.T @ before)) <= 5e-15 * np.max(np.abs(before.sum(0))) * np.abs(st.e).max()
    # fixed point
    feq = np.stack([oracle_lib.equilibrium(st, 1.1, [0.05, -0.02, 0.01][:st.D], kind=0)]).T.copy()
    fixed = feq.copy()
    oracle_lib.collide_bgk(st, fixed, 0.03, 0.1, equilibrium=0)
    if name != "D2Q25H" and name != "D3Q45":      # 2nd-order f_eq on lattices whose 2nd moments are exact
        assert np.max(np.abs(fixed - feq)) <= 1e-15


def test_equilibrium_moments_d2q9(oracle_lib):
    """Equilibrium_test (Equilibrium_test.cpp:43-109): rho, u recovered from BGK f_eq; u=(0.1,0.2)."""
    st = oracle_lib.Stencil("D2Q9", 1.0)
    feq = oracle_lib.equilibrium(st, 1.0, [0.1, 0.2], kind=0)
    assert abs(feq.sum() - 1.0) <= 1e-7
    assert np.max(np.abs(st.e.T @ feq - [0.1, 0.2])) <= 1e-7


def test_equilibrium_moments_d3q45_quartic(oracle_lib):
    """Equilibrium_D3Q45_test (:161-213): quartic f_eq with T = 1.2, u = (0.1,0.2,0.3) recovers rho, u to 1e-5 %,
    and the temperature through calculateTemperature with g = f_eq T (2 Cv - D)."""
    st = oracle_lib.Stencil("D3Q45", 1.0)
    u, T, rho, gamma = np.array([0.1, 0.2, 0.3]), 1.2, 1.0, 1.4
    feq = oracle_lib.equilibrium(st, rho, u, T=T, kind=1)
    assert abs(feq.sum() - rho) <= 1e-7
    assert np.max(np.abs(st.e.T @ feq / feq.sum() - u)) <= 1e-7
    f = feq[:, None].copy()
    g = (feq * T * (2.0 / (gamma - 1.0) - 3))[:, None].copy()
    r, v, Tout, _, rc = oracle_lib.collide_bgk_fg(st, f, g, 0.01, 0.1, equilibrium=1, gamma=gamma)
    assert rc == 0 and abs(Tout[0] - T) <= 1e-7 and abs(r[0] - rho) <= 1e-7
    assert np.max(np.abs(f[:, 0] - feq)) <= 1e-12        # equilibrium is a fixed point of relaxWithG


def test_quartic_init_matches_collision_equilibrium(oracle_lib):
    """CompressibleCFDSolver::calcQuarticEquilibrium (init, literal loop nest) == QuarticEquilibrium::polynomial."""
    for name, D in (("D2Q25H", 2), ("D3Q45", 3)):
        st = oracle_lib.Stencil(name, 1.3)
        rho, T = np.array([1.07]), np.array([1.15])
        u = np.array([[0.11], [-0.07], [0.05]])[:D] * 1.3
        f, g = fields.quartic_equilibrium_init(st.e, st.w, st.cs2, st.scaling, rho, u, T, 1.4)
        feq = oracle_lib.equilibrium(st, 1.07, u[:, 0] / 1.3, T=1.15, kind=1)
        assert np.max(np.abs(f[:, 0] - feq) / np.abs(feq)) <= 1e-12


def test_constant_streaming_and_row_structure():
    """SemiLagrangian2D/3D_ConstantStreaming_test (SemiLagrangian_test.cpp:519-596): M*1 = 1;
    2D p=3 ref 3 D2Q9 dt=0.1 / 3D p=1 ref 3 D3Q19 dt=0.1 on the unit-square/cube periodic test domain;
    plus the nnz-per-row structure of SURVEY 8 ((p+1)^k)."""
    from oracle import stencils
    for dim, p, name in ((2, 3, "D2Q9"), (3, 1, "D3Q19")):
        e, w, cs2, vm = stencils.make(name, 1.0)
        mesh = assembly.CartesianMesh.uniform(dim, 8 if dim == 2 else 4, L=1.0)
        blocks, dofs = assembly.assemble_semilagrangian(mesh, p, e, 0.1 if dim == 2 else 0.05)
        ones = np.ones(dofs.N)
        for k, m in blocks.items():
            assert k[0] == k[1]
            assert np.sum((cpu.spmv_csr(m, ones) - ones) ** 2) <= 1e-6
    o = common.oracle_problem("c1_tgv2d_d2q9")
    assert o["dofs"].N == 1089
    nnz = sum(b.nnz for b in o["blocks"].values())
    assert nnz == 120 * 1089                     # 4*5 + 4*25 per DoF
    for (bi, _), m in o["blocks"].items():
        k = np.count_nonzero(o["st"].e[bi + 1])
        assert np.all(np.diff(m.indptr) == 5 ** k)


def test_uniform_flow_stays_uniform(oracle_lib):
    """CFDSolver_SteadyStreaming_test (CFDSolver_test.cpp:44-112): |rho-1|, |u-0.1| < 1e-5 after 100 steps."""
    o = common.oracle_problem("tgv2d_small")
    st, n = o["st"], o["dofs"].N
    f = fields.equilibrium_init(st.e, st.w, st.cs2, np.ones(n), np.full((2, n), 0.1))
    stepper = oracle_lib.ReferenceOrderStepper(st, o["blocks"], n, 1.0, o["dt"])
    for _ in range(100):
        assert stepper.step(f) == 0
    assert np.max(np.abs(stepper.rho - 1)) < 1e-5 and np.max(np.abs(stepper.u - 0.1)) < 1e-5


def test_config1_energy_decay(oracle_lib):
    """Integration test #11 ConvergenceTestSemiLagrangianPeriodic (IntegrationTestCases.cpp:885-961) =
    BASELINE config 1: E_kin(t)/E_kin(0) = exp(-4 nu t) within 1e-2 at t = 0.5."""
    o = common.oracle_problem("c1_tgv2d_d2q9")
    st, n, dt, nu = o["st"], o["dofs"].N, o["dt"], 1.0
    f = o["f"].copy()
    stepper = oracle_lib.ReferenceOrderStepper(st, o["blocks"], n, nu, dt)
    r0, u0, _ = oracle_lib.collide_bgk(st, f, nu, dt)
    E0 = 0.5 * (r0 * (u0 ** 2).sum(0)).sum()
    steps = int(round(0.5 / dt))
    for _ in range(steps):
        stepper.step(f)
    E = 0.5 * (stepper.rho * (stepper.u ** 2).sum(0)).sum()
    assert abs(E / E0 - math.exp(-4 * nu * steps * dt)) < 1e-2


def test_density_exception_code(oracle_lib):
    st = oracle_lib.Stencil("D2Q9", 1.0)
    f = np.zeros((9, 4))
    assert oracle_lib.collide_bgk(st, f, 0.1, 0.1)[2] == -1      # CollisionException: rho < 1e-10


# ---- entropic family (legacy CollisionModel interface) -------------------------------------------------
def test_kbc_d2q9_mass_equals_bgk(oracle_lib):
    """KBCStandard_collideAll_test (KBCStandard_test.cpp:36-108): f_i = 0.1 i + 0.1, tau = 0.9, dt = 0.1; the sum of
    the populations after the entropic collision equals the one after BGK to 1e-5.  (The reference test instantiates
    KBCCentral; the model on the hot path is KBCStandard, same property.)"""
    st = oracle_lib.Stencil("D2Q9", 1.0)
    n, dt, tau = 9, 0.1, 0.9
    nu = tau * dt * st.cs2
    f = (0.1 * np.arange(9)[:, None] + 0.1) * np.ones((9, n))
    kbc, bgk = f.copy(), f.copy()
    _, _, rc = oracle_lib.collide_entropic(st, kbc, nu, dt, "KBC_STANDARD")
    assert rc == 0
    oracle_lib.collide_bgk(st, bgk, nu, dt)
    assert np.max(np.abs(kbc.sum(0) - bgk.sum(0))) <= 1e-5
    assert np.max(np.abs(kbc.sum(0) - f.sum(0))) <= 1e-14 * f.sum(0).max()
    assert np.max(np.abs(kbc - bgk)) > 1e-3          # and it is not simply BGK


def test_kbc_d3q15_mass_equals_bgk(oracle_lib):
    """KBCStandard_collideAllD3Q15_test (KBCStandard_test.cpp:111-189), same inputs: synthetic populations, tau 0.9."""
    st = oracle_lib.Stencil("D3Q15", 1.0)
    n, dt, tau = 27, 0.1, 0.9
    nu = tau * dt * st.cs2
    f = synthetic_populations(15, n)
    kbc, bgk = f.copy(), f.copy()
    rho, u, rc = oracle_lib.collide_entropic(st, kbc, nu, dt, "KBC_STANDARD")
    assert rc == 0
    oracle_lib.collide_bgk(st, bgk, nu, dt)
    assert np.max(np.abs(kbc.sum(0) - bgk.sum(0))) <= 1e-5
    assert np.max(np.abs(rho - f.sum(0))) <= 1e-14 * rho.max()
    # momentum is a collision invariant too (scaling 1, where the reference's u^2 quirk is harmless)
    assert np.max(np.abs(st.e.T @ kbc - st.e.T @ f)) <= 1e-13 * f.sum(0).max()


def test_kbc_equilibrium_is_fixed_point(oracle_lib):
    """The product-form entropic equilibrium (KBCStandard.cpp:250-271) is reproduced by the collision (gamma = 2 branch
    for sum_h < 1e-16, :430-433) and carries the prescribed rho, u."""
    st = oracle_lib.Stencil("D2Q9", 1.0)
    rho0, u0 = 1.07, np.array([0.04, -0.03])
    r3 = math.sqrt(3.0)
    w = st.w
    feq = np.empty(9)
    for i in range(9):
        val = w[i] * rho0
        for p in range(2):
            v = u0[p] * r3
            val *= 2 - math.sqrt(1 + v * v)
            val *= ((2 * v / r3 + math.sqrt(1 + v * v)) / (1 - v / r3)) ** st.e[i, p]
        feq[i] = val                                   # KBCStandard::getEquilibriumDistribution, :26-62
    f = np.repeat(feq[:, None], 4, axis=1).copy()
    rho, u, rc = oracle_lib.collide_entropic(st, f, 0.01, 0.1, "KBC_STANDARD")
    assert rc == 0
    assert np.max(np.abs(f - feq[:, None])) <= 1e-15
    assert np.max(np.abs(rho - rho0)) <= 1e-15 and np.max(np.abs(u - u0[:, None])) <= 1e-15


def test_mrt_entropic_d3q19_invariants_and_tables(oracle_lib):
    """MRTEntropic::collideAllD3Q19 (MRTEntropic.cpp:167-305): invm is the inverse of tm; density and the momentum
    moments m3, m5, m7 (as the reference defines them through tm) are untouched; the density guard looks at the density
    of the previous call."""
    tm, invm = oracle_lib.mrt_entropic_tables()
    assert np.max(np.abs(invm @ tm - np.eye(19))) <= 1e-15
    st = oracle_lib.Stencil("D3Q19", 1.0)
    f = synthetic_populations(19, 30) * st.w[:, None]
    g = f.copy()
    rho, u, rc = oracle_lib.collide_entropic(st, g, 0.02, 0.1, "MRT_ENTROPIC")
    assert rc == 0
    for row in (0, 3, 5, 7):
        assert np.max(np.abs(tm[row] @ g - tm[row] @ f)) <= 1e-14 * np.abs(f).sum(0).max()
    assert np.max(np.abs(rho - f.sum(0))) <= 1e-15 * rho.max()
    assert np.max(np.abs(u[0] - (tm[3] @ f) / rho)) <= 1e-15
    g = f.copy()
    _, _, rc = oracle_lib.collide_entropic(st, g, 0.02, 0.1, "MRT_ENTROPIC", rho_prev=np.zeros(30))
    assert rc == 1 and np.array_equal(g, f)


def test_entropic_dispatch_matches_reference(oracle_lib):
    """KBCStandard::collideAll throws for anything but D2Q9 / D3Q15 (KBCStandard.cpp:70-85); MRTEntropic for
    anything but D2Q9 / D3Q19 (MRTEntropic.cpp:27-36; only D3Q19 is on the path)."""
    with pytest.raises(ValueError):
        oracle_lib.collide_entropic(oracle_lib.Stencil("D3Q19", 1.0), np.ones((19, 4)), 0.1, 0.1, "KBC_STANDARD")
    with pytest.raises(ValueError):
        oracle_lib.collide_entropic(oracle_lib.Stencil("D3Q15", 1.0), np.ones((15, 4)), 0.1, 0.1, "MRT_ENTROPIC")


# ---------------------------------------------------------------------------------------------
# collision_advanced Regularized / MultipleRelaxationTime / forcing
# ---------------------------------------------------------------------------------------------
def _relaxation_test_f(Q):
    """Relaxation_test.cpp:37-40: f_i = 1.5 + sin(1.5 i) + 0.001 + i/(i+1) + (0.5 cos 0.5)^2, integer i/(i+1) = 0."""
    i = np.arange(Q, dtype=np.float64)
    return (1.5 + np.sin(1.5 * i) + 0.001 + (0.5 * math.cos(0.5)) ** 2)[:, None].copy()


def test_mrt_tables_golden_vs_product():
    """The product's host mirror rebuilds the three MRT bases from their definitions; the oracle uses the reference's
    literals (tests/golden/mrt_tables.npz, AuxiliaryMRTFunctions.cpp:15-205).  They agree to round-off, M T = I,
    and make_diag matches entry by entry."""
    from natrium_b200 import mrt
    for pb, ob in [(mrt.DELLAR_D2Q9, "DELLAR_D2Q9"), (mrt.LALLEMAND_D2Q9, "LALLEMAND_D2Q9"), (mrt.DHUMIERES_D3Q19, "DHUMIERES_D3Q19")]:
        M, T = cpu.mrt_tables(ob)
        assert np.array_equal(mrt.make_M(pb), M)
        assert np.max(np.abs(mrt.make_T(pb) - T)) <= 1e-16
        assert np.max(np.abs(M @ T - np.eye(len(M)))) <= 1e-15
        for pm, om in [(mrt.RELAX_FULL, "RELAX_FULL")] + ([(mrt.DELLAR_RELAX_ONLY_N, "DELLAR_RELAX_ONLY_N")] if ob == "DELLAR_D2Q9" else []) \
                + ([(mrt.RELAX_DHUMIERES_PAPER, "RELAX_DHUMIERES_PAPER")] if ob == "DHUMIERES_D3Q19" else []):
            assert np.array_equal(mrt.make_diag(0.8, pb, pm), cpu.mrt_diag(0.8, ob, om))


def test_relaxation_regularized(oracle_lib):
    """Relaxation_Regularized_test (Relaxation_test.cpp:30-68): the x-velocity recomputed from the relaxed populations
    stays what calculateVelocity gives (the test's bound is 1e-5; momentum is conserved to round-off)."""
    st = oracle_lib.Stencil("D2Q9", 1.0)
    f = _relaxation_test_f(9)
    before = f.copy()
    rho, u, rc = oracle_lib.collide_advanced(st, f, 1.0, 0.1, scheme="BGK_REGULARIZED")
    assert rc == 0
    ux = (f[1] - f[3] + f[5] - f[6] - f[7] + f[8]) / rho
    assert abs(ux[0] - u[0, 0]) < 1e-5
    assert abs(f.sum() - before.sum()) <= 1e-13 * before.sum()
    assert np.max(np.abs(st.e.T @ f - st.e.T @ before)) <= 1e-13 * before.sum()


@pytest.mark.parametrize("name,basis", [("D2Q9", "DELLAR_D2Q9"), ("D2Q9", "LALLEMAND_D2Q9"), ("D3Q19", "DHUMIERES_D3Q19")])
def test_relaxation_mrt_conserves(name, basis, oracle_lib):
    """Relaxation_MRT_D2Q9_test / _D3Q19_test (:74-196): density and velocity unchanged to 1e-10 % by the MRT
    relaxation; scaling 4, dt 0.1, viscosity 1."""
    st = oracle_lib.Stencil(name, 4.0)
    f = _relaxation_test_f(st.Q)
    rho0 = f.sum(0)
    u0 = (st.e / st.scaling).T @ f / rho0
    _, _, rc = oracle_lib.collide_advanced(st, f, 1.0, 0.1, scheme="MRT_STANDARD", mrt_basis=basis)
    assert rc == 0
    rho1 = f.sum(0)
    u1 = (st.e / st.scaling).T @ f / rho1
    assert abs(rho1[0] - rho0[0]) <= 1e-12 * rho0[0]
    assert np.max(np.abs(u1 - u0)) <= 1e-12 * np.max(np.abs(u0))


def test_relaxation_equiv_mrt_regularized(oracle_lib):
    """Relaxation_Equiv_MRT_Reg_test (:200-255): Dellar-D2Q9 MRT with full relaxation of the ghost moments equals the
    regularized scheme to 1e-10 %."""
    st = oracle_lib.Stencil("D2Q9", 4.0)
    f1, f2 = _relaxation_test_f(9), _relaxation_test_f(9)
    oracle_lib.collide_advanced(st, f1, 1.0, 0.1, scheme="MRT_STANDARD", mrt_basis="DELLAR_D2Q9")
    oracle_lib.collide_advanced(st, f2, 1.0, 0.1, scheme="BGK_REGULARIZED")
    assert np.max(np.abs(f1 - f2) / np.abs(f2)) <= 1e-12


def test_forcing_hooks(oracle_lib):
    """External-force hooks of collideAll (CollisionOperator.h:79-96; Aux...h:332-417): without a force the advanced
    entry equals the BGK oracle bit for bit; both force types add dt*F to the momentum to round-off (shifting
    velocity: through the shifted equilibrium, tau*(1/tau); exact difference: f_eq(u + dt F/rho) - f_eq(u)); the stored
    velocity is shifted by dt*F/(2 rho); NO_FORCING with a force and GUO raise (status -2 / -3)."""
    st = oracle_lib.Stencil("D3Q19", 2.0)
    n = 40
    f0 = synthetic_populations(st.Q, n) * st.w[:, None]
    nu, dt, F = 0.05, 0.1, np.array([1e-3, -2e-3, 5e-4])
    a, b = f0.copy(), f0.copy()
    ra, ua, _ = oracle_lib.collide_bgk(st, a, nu, dt)
    rb, ub, rc = oracle_lib.collide_advanced(st, b, nu, dt)
    assert rc == 0 and np.array_equal(a, b) and np.array_equal(ua, ub)
    mom0 = st.e.T @ f0
    for ft in ["SHIFTING_VELOCITY", "EXACT_DIFFERENCE"]:
        c = f0.copy()
        rho, u, rc = oracle_lib.collide_advanced(st, c, nu, dt, force=F, force_type=ft)
        assert rc == 0
        dmom = st.e.T @ c - mom0
        assert np.max(np.abs(dmom - dt * F[:, None] * np.ones(n))) <= 1e-13
        assert np.max(np.abs(u - (ua + 0.5 * dt * F[:, None] / rho))) <= 1e-14
        assert np.max(np.abs(c.sum(0) - f0.sum(0))) <= 1e-14
    assert oracle_lib.collide_advanced(st, f0.copy(), nu, dt, force=F, force_type="NO_FORCING")[2] == -2
    assert oracle_lib.collide_advanced(st, f0.copy(), nu, dt, force=F, force_type="GUO")[2] == -3


# ---------------------------------------------------------------------------------------------
# wall hits
# ---------------------------------------------------------------------------------------------
def test_wall_hits_properties(oracle_lib):
    """ThermalBounceBack (ThermalBounceBack.cpp:50-109) re-equilibrates the destination DoF to the wall temperature:
    afterwards density and velocity of the DoF are unchanged, its temperature is T_wall up to the energy of the
    non-equilibrium part that is kept, and g = g_eq(T_w).  VelocityNeqBounceBack adds its term to one population."""
    from natrium_b200 import harness
    from natrium_b200.stencils import Stencil as PStencil
    st = oracle_lib.Stencil("D3Q45", 1.0)
    pst = PStencil("D3Q45", 1.0)
    n, gamma, Tw = 12, 1.4, 0.85
    rng = np.random.default_rng(3)
    rho = 1.0 + 0.05 * rng.standard_normal(n)
    u = 0.05 * rng.standard_normal((3, n))
    T = 1.0 + 0.05 * rng.standard_normal(n)
    f, g = harness.quartic_equilibrium_distributions(pst, rho, u, T, gamma)
    f *= 1.0 + 0.01 * rng.standard_normal(f.shape)
    f0, g0 = f.copy(), g.copy()
    idx = np.array([3, 7, 7, 1], dtype=np.int32)
    dirs = np.array([5, 9, 11, 2], dtype=np.int32)
    kinds = np.array([1, 1, 1, 0], dtype=np.int32)
    vals = np.array([Tw, Tw, Tw, 0.125])
    assert oracle_lib.apply_wall_hits(st, f, g, idx, dirs, kinds, vals) == 0
    e = st.e
    for i in (3, 7):
        r0, r1 = f0[:, i].sum(), f[:, i].sum()
        assert abs(r1 - r0) <= 1e-13
        assert np.max(np.abs(e.T @ f[:, i] - e.T @ f0[:, i])) <= 1e-12
        ui = e.T @ f[:, i] / r1
        Cv = 1.0 / (gamma - 1.0)
        Ti = 0.5 * (np.sum(((e - ui) ** 2).sum(1) * f[:, i]) / st.cs2 + g[:, i].sum()) / (r1 * Cv)
        assert abs(Ti - Tw) <= 1e-3           # up to the energy carried by the kept non-equilibrium part
    untouched = [k for k in range(n) if k not in (1, 3, 7)]
    assert np.array_equal(f[:, untouched], f0[:, untouched]) and np.array_equal(g[:, untouched], g0[:, untouched])
    assert f[2, 1] == f0[2, 1] + 0.125 and np.array_equal(np.delete(f[:, 1], 2), np.delete(f0[:, 1], 2))
    assert np.array_equal(g[:, 1], g0[:, 1])
    # thermal hits need D3Q45 and g
    st19 = oracle_lib.Stencil("D3Q19", 1.0)
    assert oracle_lib.apply_wall_hits(st19, np.ones((19, 4)), None, [0], [1], [1], [0.85]) == -3


# ---------------------------------------------------------------------------------------------
# pseudo-entropic stabilizer
# ---------------------------------------------------------------------------------------------
def test_stabilizer_tables_and_properties(oracle_lib):
    """The host mirror rebuilds the three stabilizer matrices from their moment-space definition; they equal the
    reference's literals (PseudoEntropicStabilizer.cpp:27-150) to round-off.  The matrices are projections that
    keep density, momentum and the second moments, and leave a BGK equilibrium at rest unchanged."""
    from natrium_b200 import mrt
    for name, with_e in [("D2Q9", False), ("D2Q9", True), ("D3Q19", False)]:
        A = oracle_lib.stabilizer_matrix(name, with_e)
        assert np.max(np.abs(mrt.make_stabilizer(name, with_e) - A)) <= 2e-16
        assert np.max(np.abs(A @ A - A)) <= 1e-15
        st = oracle_lib.Stencil(name, 1.0)
        f = synthetic_populations(st.Q, 7) * st.w[:, None]
        g = f.copy()
        oracle_lib.apply_stabilizer(g, A)
        assert np.max(np.abs(g.sum(0) - f.sum(0))) <= 1e-15
        assert np.max(np.abs(st.e.T @ g - st.e.T @ f)) <= 1e-15
        if not with_e:     # the with-e variant also replaces the energy moment
            P0 = np.einsum("qa,qb,qn->abn", st.e, st.e, f)
            P1 = np.einsum("qa,qb,qn->abn", st.e, st.e, g)
            assert np.max(np.abs(P1 - P0)) <= 1e-15
        rest = st.w[:, None] * np.ones((st.Q, 1))
        r2 = rest.copy()
        oracle_lib.apply_stabilizer(r2, A)
        assert np.max(np.abs(r2 - rest)) <= 1e-15


def test_config1_energy_decay_with_stabilizer(oracle_lib):
    """Integration test #11 as the reference runs it: with the PseudoEntropicStabilizer appended as data processor
    (IntegrationTestCases.cpp:924); same bound, E_kin(t)/E_kin(0) = exp(-2) +- 1e-2 at t = 1/(2 nu)."""
    o = common.oracle_problem("c1_tgv2d_d2q9")
    st, n, dt, nu = o["st"], o["dofs"].N, o["dt"], 1.0
    A = oracle_lib.stabilizer_matrix("D2Q9")
    f = o["f"].copy()
    stepper = oracle_lib.ReferenceOrderStepper(st, o["blocks"], n, nu, dt)
    r0, u0, _ = oracle_lib.collide_bgk(st, f, nu, dt)
    E0 = 0.5 * (r0 * (u0 ** 2).sum(0)).sum()
    steps = int(round(0.5 / dt))
    for _ in range(steps):
        stepper.step(f)
        oracle_lib.apply_stabilizer(f, A)
    E = 0.5 * (stepper.rho * (stepper.u ** 2).sum(0)).sum()
    assert abs(E / E0 - math.exp(-4 * nu * steps * dt)) < 1e-2


def test_walled_assembly_bounce_blocks_and_hits(oracle_lib):
    """Wall treatment of fillSparseObject (SemiLagrangian.cpp:358-384): a path that reaches a VelocityNeqBounceBack /
    ThermalBounceBack wall is reflected with the opposite direction, so its row lands in the off-diagonal block
    (alpha, opposite(alpha)), rows still sum to one (the property SemiLagrangian*_ConstantStreaming_test pins), and
    one BoundaryHit per bounced path is recorded with 0 <= dtHit <= dt (BoundaryHit.h:40-56)."""
    st = oracle_lib.Stencil("D2Q9", 3.0)
    opp = np.array([int(np.argmin(np.abs(st.e + st.e[i]).sum(1))) for i in range(9)])
    mesh = assembly.CartesianMesh([np.linspace(0, 2.0, 6), 2.0 * np.array([0, 0.2, 0.45, 0.75, 1.0])], boundary=["periodic", "wall"])
    dt = assembly.calculate_timestep(mesh, 3, st.max_speed, 0.8)
    blocks, dofs = assembly.assemble_semilagrangian(mesh, 3, st.e, dt, opposite=opp)
    ones = np.ones(dofs.N)
    for a in range(8):
        y = sum(blocks[k] @ ones for k in blocks if k[0] == a)
        assert np.max(np.abs(y - 1.0)) <= 1e-12
    assert {k for k in blocks if k[0] != k[1]} <= {(a - 1, opp[a] - 1) for a in range(1, 9)}
    assert len(dofs.hits) > 0
    for h in dofs.hits:
        a = h["direction"]
        assert st.e[a][1] != 0                                   # only directions with a wall-normal component bounce
        assert blocks[(a - 1, opp[a] - 1)][h["index"]].nnz > 0
        assert -1e-14 <= h["dt_hit"] <= dt + 1e-14
    keys = [(h["cell"], h["point"]) for h in dofs.hits]
    assert keys == sorted(keys)                                  # HitList iteration order: cell, then point


def test_face_crossed_first_kats():
    """SemiLagrangian2D/3D_FaceCrossedFirst_test (test/advection/SemiLagrangian_test.cpp:227-456): rays from the
    barycentre (0.0625, ...) of the lower-left cell of the 8^dim unit-square / unit-cube test domain; expected face id,
    lambda and boundary point are the reference's literals."""
    for dim in (2, 3):
        mesh = assembly.CartesianMesh.uniform(dim, 8, L=1.0)
        cell = (0,) * dim
        c = [0.0625] * dim

        def ray(**kw):
            p2 = list(c)
            for k, v in kw.items():
                p2["xyz".index(k)] = v
            return assembly.face_crossed_first(mesh, cell, c, p2)

        assert ray()[0] == -1                                        # no face crossed
        kats = [(dict(x=-0.0625), 0, 0.5, {0: 0.0}), (dict(x=0.1875), 1, 0.5, {0: 0.125}),
                (dict(y=-0.0625), 2, 0.5, {1: 0.0}), (dict(y=0.1875), 3, 0.5, {1: 0.125})]
        if dim == 3:
            kats += [(dict(z=-0.0625), 4, 0.5, {2: 0.0}), (dict(z=0.1875), 5, 0.5, {2: 0.125}),
                     # three faces crossed, the one hit first wins (:419-452)
                     (dict(x=0.0625 - 0.25, y=-0.0625, z=-0.0625), 0, 0.25, {0: 0.0, 1: 0.03125, 2: 0.03125}),
                     (dict(x=-0.0625, y=0.0625 - 0.25, z=-0.0625), 2, 0.25, {0: 0.03125, 1: 0.0, 2: 0.03125}),
                     (dict(x=-0.0625, y=-0.0625, z=0.0625 - 0.25), 4, 0.25, {0: 0.03125, 1: 0.03125, 2: 0.0})]
        else:
            kats += [(dict(x=0.0625 - 0.25, y=-0.0625), 0, 0.25, {0: 0.0, 1: 0.03125}),     # faces 0 and 2, 0 first (:283-290)
                     (dict(x=-0.0625, y=0.0625 - 0.25), 2, 0.25, {0: 0.03125, 1: 0.0})]     # faces 0 and 2, 2 first (:291-299)
        for kw, face, lam, pb_expect in kats:
            f, pb, l = ray(**kw)
            assert f == face, (dim, kw, f)
            assert abs(l - lam) <= 1e-13
            for d in range(dim):
                assert abs(pb[d] - pb_expect.get(d, 0.0625)) <= 1e-13, (dim, kw, pb)


@pytest.mark.parametrize("scaling", [1.0, 5.0])
def test_bgk_equilibrium_moments(scaling, oracle_lib):
    """BGKMoments_test / D2Q9IncompressibleModelMoments_Scaled_test (test/collision/BGKStandard_test.cpp:69-152,189-263):
    rho = 1.45, u = (2.3, -1.14); density, momentum and momentum-flux tensor rho u u + rho cs^2 I of the legacy
    equilibrium to 1e-14 (the reference's TOLERANCE), unscaled and with stencil scaling 5; and the collision_advanced
    BGKEquilibrium (unscaled directions, u / scaling) is the same distribution."""
    st = oracle_lib.Stencil("D2Q9", scaling)
    rho, u = 1.45, np.array([2.3, -1.14])
    feq = oracle_lib.legacy_feq(st, rho, u)
    tol = 1e-14 * max(1.0, scaling * scaling)          # the reference's absolute 1e-14 at scaling 1; entries grow with scaling^2
    assert abs(feq.sum() - rho) <= 1e-14
    assert np.max(np.abs(st.e.T @ feq - rho * u)) <= tol
    P = np.einsum("qa,qb,q->ab", st.e, st.e, feq)
    assert np.max(np.abs(P - (rho * np.outer(u, u) + rho * st.cs2 * np.eye(2)))) <= tol
    adv = oracle_lib.equilibrium(st, rho, u / scaling, kind=0)
    assert np.max(np.abs(adv - feq)) <= 1e-14


def test_poiseuille_bounce_back_with_forcing(oracle_lib):
    """SemiLagrangianBoundaryHandler_PoiseuilleBB_test (test/boundaries/SemiLagrangianBoundaryHandler_test.cpp:175-229):
    body-force driven channel between two VelocityNeqBounceBack walls with the SHIFTING_VELOCITY forcing scheme;
    the converged mean x-velocity lies within 10 % of u_bulk.  Walls (bounce blocks + hit list), forcing and the
    collision all come from the oracle's restatements."""
    pb = common.poiseuille_problem(oracle_lib)
    st, f = pb["st"], pb["f0"].copy()
    idx, dirs, kinds, vals = pb["hits"]
    mean_prev, u = None, None
    for it in range(6000):
        f = oracle_lib.stream(pb["blocks"], f)
        assert oracle_lib.apply_wall_hits(st, f, None, idx, dirs, kinds, vals) == 0
        _, u, rc = oracle_lib.collide_advanced(st, f, pb["nu"], pb["dt"], force=pb["F"], force_type="SHIFTING_VELOCITY")
        assert rc == 0
        if it % 100 == 99:
            mean = common.integral_mean(pb["dofs"], pb["mesh"], 2, u[0])
            if mean_prev is not None and abs(mean - mean_prev) <= 1e-6 * abs(mean):     # setConvergenceThreshold(1e-6)
                break
            mean_prev = mean
    mean = common.integral_mean(pb["dofs"], pb["mesh"], 2, u[0])          # SolverStats::getMeanVelocityX
    assert 0.9 * pb["u_bulk"] < mean < 1.1 * pb["u_bulk"], (mean, pb["u_bulk"], it)
    assert abs(common.integral_mean(pb["dofs"], pb["mesh"], 2, u[1])) < 1e-3 * pb["u_bulk"]


def test_moving_walls_drag_the_fluid(oracle_lib):
    """VelocityNeqBounceBack2D_SL_BoundaryVelocity_test / _MassConservation_test
    (test/boundaries/VelocityNeqBounceBack_SL_test.cpp:49-87 with WallFixture.h): after 100 steps every DoF moves with
    the walls, max |u - 0.01| < 1e-3 and max |v| < 1e-5, and the mean density stays 1 to 1e-10."""
    pb = common.moving_walls_problem(oracle_lib)
    st, f = pb["st"], pb["f0"].copy()
    idx, dirs, kinds, vals = pb["hits"]
    assert np.count_nonzero(vals) > 0
    for _ in range(100):
        f = oracle_lib.stream(pb["blocks"], f)
        assert oracle_lib.apply_wall_hits(st, f, None, idx, dirs, kinds, vals) == 0
        rho, u, rc = oracle_lib.collide_bgk(st, f, pb["nu"], pb["dt"])
        assert rc == 0
    assert np.max(np.abs(u[0] - 0.01)) < 1e-3
    assert np.max(np.abs(u[1])) < 1e-5
    assert abs(np.abs(rho).sum() / pb["dofs"].N - 1.0) < 1e-10


# ---------------------------------------------------------------------------------------------------
# ExponentialFilter: set-up restated in oracle/filter.py, cell loop in oracle/natrium_oracle.c
# ---------------------------------------------------------------------------------------------------
def test_exponential_filter_projection_kat():
    """ExponentialFilter_TestProjection_test (test/smoothing/ExponentialFilter_test.cpp:80-116): p = 4, 1-d, nodal
    coefficients 0.5 .. 0.9; the Legendre expansion built from getProjectToLegendre() takes the same value at x = 0.3 as the
    nodal expansion (BOOST_CHECK_CLOSE 1e-10 %)."""
    from oracle import filter as F
    p = 4
    to, fr = F.projection_matrices(p, 1)
    dgq = np.array([0.5, 0.6, 0.7, 0.8, 0.9])
    nodes, w = F.gauss_lobatto_01(p + 1)
    assert abs(w.sum() - 1.0) < 1e-15
    expected = float(dgq @ F.lagrange_1d(nodes, 0.3))
    leg = to @ dgq
    result = sum(leg[i] * F.legendre_01(i, 0.3) for i in range(p + 1))
    assert abs(result - expected) <= 1e-12 * abs(expected)
    assert np.abs(to @ fr - np.eye(p + 1)).max() < 1e-13
    # ExponentialFilter_PolynomialDegree_test (:30-46): mode k is a polynomial of degree k, orthonormal on [0, 1]
    xq, wq = np.polynomial.legendre.leggauss(12)
    xq, wq = 0.5 * (xq + 1), 0.5 * wq
    G = np.array([[np.sum(wq * F.legendre_01(i, xq) * F.legendre_01(j, xq)) for j in range(6)] for i in range(6)])
    assert np.abs(G - np.eye(6)).max() < 1e-13


@pytest.mark.parametrize("p,dim", [(4, 1), (3, 2), (2, 3), (4, 3)])
def test_exponential_filter_host_mirror_equals_oracle_setup(p, dim):
    """natrium_b200.host.ExponentialFilter (from_legendre = mode values at the nodes, to_legendre = its inverse) against the
    oracle's literal quadrature sums; damping factors for degree-by-maximum and degree-by-sum, and the reference's 3-d
    degree quirk (iy = iz)."""
    from oracle import filter as F
    from natrium_b200 import host
    to, fr = F.projection_matrices(p, dim)
    for by_sum, alpha, s, Nc in [(False, 36.0, 2.0, 1), (True, 10.0, 4.0, 2)]:
        h = host.ExponentialFilter(alpha, s, Nc, by_sum, p, dim)
        assert np.abs(h.getProjectToLegendre() - to).max() <= 1e-12 * np.abs(to).max()
        assert np.abs(h.getProjectFromLegendre() - fr).max() <= 1e-12 * np.abs(fr).max()
        sg, damped = F.damping(p, dim, alpha, s, Nc, by_sum)
        assert np.array_equal(h.sigma, sg)
        assert np.all(sg[damped == 0] == 1.0) and np.all(sg[damped == 1] < 1.0)
    if dim == 3:
        dq, _ = F.degree_vectors(p, 3, reference_quirk=True)
        di, _ = F.degree_vectors(p, 3, reference_quirk=False)
        assert not np.array_equal(dq, di)             # mode (0, 1, 0): the reference sees degree 0
        assert dq[(p + 1)] == 0 and di[(p + 1)] == 1


def test_exponential_filter_cell_loop(oracle_lib):
    """applyFilter (ExponentialFilter.cpp:139-199) on a periodic 2-d mesh: the C loop equals a plain numpy restatement of the
    sequential cell loop; constants and the element-wise mean mode are untouched; the result depends on the cell order
    (shared face DoFs are read after earlier cells wrote them), which is why the device must keep the order; harness and
    oracle agree on cell->get_dof_indices."""
    from oracle import assembly, filter as F
    from natrium_b200 import harness
    from natrium_b200.stencils import Stencil
    p, dim, cells = 3, 2, [4, 3]
    mesh = assembly.CartesianMesh.uniform(dim, cells)
    dofs = assembly.DofMap(mesh, p)
    st = Stencil("D2Q9", 3.0)
    pb = harness.CartesianProblem(dim, cells, p)
    part = harness.SlabPartition(pb, st, pb.timestep(st, 0.4))
    cd = part.cell_dofs()
    assert np.array_equal(cd, np.array([dofs.cell_dofs(c) for c in mesh.cells()], dtype=np.int32))
    to, fr = F.projection_matrices(p, dim)
    sg, damped = F.damping(p, dim, 36.0, 2.0, 1)
    rng = np.random.default_rng(3)
    v0 = rng.standard_normal(pb.N)
    v = F.apply_filter(cd, to, fr, sg, damped, v0.copy())
    w = v0.copy()
    for c in range(cd.shape[0]):
        leg = to @ w[cd[c]]
        leg = np.where(damped == 1, sg * leg, leg)
        w[cd[c]] = fr @ leg
    assert np.abs(v - w).max() <= 1e-13 * np.abs(w).max()
    ones = F.apply_filter(cd, to, fr, sg, damped, np.ones(pb.N))
    assert np.abs(ones - 1.0).max() < 1e-13
    v_rev = F.apply_filter(cd[::-1].copy(), to, fr, sg, damped, v0.copy())
    assert np.abs(v_rev - v).max() > 1e-6 * np.abs(v).max()
    # only the highest mode damped (Nc = p): a smooth field is changed little, a rough one a lot (what the filter is for)
    sg_top, damped_top = F.damping(p, dim, 36.0, 2.0, p)
    x = part.owned_points()
    smooth = np.cos(x[:, 0]) * np.sin(x[:, 1])
    fs = F.apply_filter(cd, to, fr, sg_top, damped_top, smooth.copy())
    vr = F.apply_filter(cd, to, fr, sg_top, damped_top, v0.copy())
    assert np.abs(fs - smooth).max() < 0.2 and np.abs(vr - v0).max() > 5 * np.abs(fs - smooth).max()
