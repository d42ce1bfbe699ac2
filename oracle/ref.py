"""ctypes front-end of oracle/_ref/libnatrium_ref.so: the REFERENCE's own collision and stencil code, compiled from
/root/reference by oracle/Makefile.ref (stand-ins for deal.II/Trilinos/Boost/MPI in oracle/ref_stubs/).

Test infrastructure only.  Used to pin the oracle restatement (oracle/natrium_oracle.c, entropic_oracle.c, stencils.py)
to the reference itself: tests/test_oracle_vs_ref.py.  On a machine without /root/reference (the GPU box) the
prebuilt library that travelled with the snapshot is used; `available()` says whether there is one.
"""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libnatrium_ref.so")
REFERENCE = "/root/reference/src/library/natrium"
_LIB = None
_dp = C.POINTER(C.c_double)

# ConfigNames.h:46-62,63-69,20-33,123-128 -- the reference's own enumerator values
SCHEMES = {"BGK_STANDARD": 0, "MRT_STANDARD": 5, "MRT_ENTROPIC": 6, "KBC_STANDARD": 8, "BGK_REGULARIZED": 12}
EQUILIBRIA = {"BGK_EQUILIBRIUM": 0, "QUARTIC_EQUILIBRIUM": 1}
FORCE_TYPES = {"NO_FORCING": 0, "SHIFTING_VELOCITY": 1, "EXACT_DIFFERENCE": 2, "GUO": 3}
RELAX_MODES = {"RELAX_FULL": 0, "DELLAR_RELAX_ONLY_N": 1, "RELAX_DHUMIERES_PAPER": 2}
BASES = {"DELLAR_D2Q9": 0, "LALLEMAND_D2Q9": 1, "DHUMIERES_D3Q19": 2}
STATUS = {0: "ok", -1: "CollisionException", -2: "NATriuMException", -3: "NotImplementedException", -4: "DensityZeroException"}


def build(force=False):
    """Compiles the reference sources when they are present; otherwise keeps a prebuilt library."""
    if os.path.isdir(REFERENCE):
        subprocess.run(["make", "-C", _HERE, "-s", "-f", "Makefile.ref"] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)
    return _SO if os.path.exists(_SO) else None


def available():
    return build() is not None


def lib():
    global _LIB
    if _LIB is None:
        so = build()
        if so is None:
            raise RuntimeError("oracle/_ref/libnatrium_ref.so is missing and /root/reference is not present to build it")
        _LIB = C.CDLL(so)
    return _LIB


def _d(a):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(_dp)


def stencil(name, scaling=1.0):
    """(e[Q,D] scaled, w[Q], cs2_scaled, max_speed, opposite[Q]) from the reference's stencil classes."""
    D, Q = C.c_int(), C.c_int()
    e, w, opp = np.zeros(45 * 3), np.zeros(45), np.zeros(45, dtype=np.int32)
    cs2, mx = C.c_double(), C.c_double()
    rc = lib().ref_stencil(name.encode(), C.c_double(scaling), C.byref(D), C.byref(Q), _d(e), _d(w), C.byref(cs2),
                           C.byref(mx), opp.ctypes.data_as(C.POINTER(C.c_int32)))
    if rc:
        raise ValueError(name)
    return e[:Q.value * D.value].reshape(Q.value, D.value).copy(), w[:Q.value].copy(), cs2.value, mx.value, opp[:Q.value].copy()


def select_collision(name, scaling, f, viscosity, dt, scheme="BGK_STANDARD", equilibrium="BGK_EQUILIBRIUM", g=None,
                     in_init=False, u_init=None, force=None, force_type="NO_FORCING", mrt_basis="DELLAR_D2Q9",
                     relax_mode="RELAX_FULL", gamma=1.4, prandtl=None, sutherland=False, n=None):
    """natrium::selectCollision<dim> in place on f (and g).  Returns dict(rho, u, T, sensor, status, message)."""
    Q, stride = f.shape
    n = stride if n is None else n
    D = 2 if name.startswith("D2") else 3
    rho = np.zeros(n)
    u = np.zeros((D, n)) if u_init is None else np.ascontiguousarray(u_init, dtype=np.float64).copy()
    T, mss = np.zeros(n), np.zeros(n)
    fv = np.zeros(3)
    if force is not None:
        fv[:len(force)] = force
    err = C.create_string_buffer(512)
    rc = lib().ref_select_collision(
        name.encode(), C.c_double(scaling), C.c_int(SCHEMES[scheme]), C.c_int(EQUILIBRIA[equilibrium]),
        C.c_int(FORCE_TYPES[force_type]), C.c_int(0 if force is None else 1), _d(fv), C.c_int(BASES[mrt_basis]),
        C.c_int(RELAX_MODES[relax_mode]), C.c_double(viscosity), C.c_double(dt), C.c_int(1 if in_init else 0),
        C.c_int(0 if g is None else 1), C.c_double(gamma), C.c_int(0 if prandtl is None else 1),
        C.c_double(1.0 if prandtl is None else prandtl), C.c_int(1 if sutherland else 0), C.c_int64(n), C.c_int64(stride),
        _d(f), None if g is None else _d(g), _d(rho), _d(u), _d(T), _d(mss), err, C.c_int(512))
    return dict(rho=rho, u=u, T=T, sensor=mss, status=rc, message=err.value.decode())


def legacy_collide(name, scaling, model, f, viscosity, dt, in_init=False, u_init=None, rho_prev=None, n=None):
    """CollisionModel::collideAll of the legacy family (KBC_STANDARD, MRT_ENTROPIC, MRT_STANDARD, BGK_STANDARD), in place."""
    Q, stride = f.shape
    n = stride if n is None else n
    D = 2 if name.startswith("D2") else 3
    rho = np.ones(n) if rho_prev is None else np.ascontiguousarray(rho_prev, dtype=np.float64).copy()
    u = np.zeros((D, n)) if u_init is None else np.ascontiguousarray(u_init, dtype=np.float64).copy()
    err = C.create_string_buffer(512)
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:      # KBCStandard opens deviation_KBC_STANDARD.txt in the cwd (KBCStandard.cpp:17-18)
        os.chdir(tmp)
        try:
            rc = lib().ref_legacy_collide(name.encode(), C.c_double(scaling), model.encode(), C.c_double(viscosity),
                                          C.c_double(dt), C.c_int(1 if in_init else 0), C.c_int64(n), C.c_int64(stride),
                                          _d(f), _d(rho), _d(u), err, C.c_int(512))
        finally:
            os.chdir(cwd)
    return dict(rho=rho, u=u, status=rc, message=err.value.decode())


def legacy_feq(name, scaling, rho, u_scaled):
    e, _, _, _, _ = stencil(name, scaling)
    feq = np.zeros(e.shape[0])
    uu = np.ascontiguousarray(u_scaled, dtype=np.float64)
    lib().ref_legacy_feq(name.encode(), C.c_double(scaling), C.c_double(rho), _d(uu), _d(feq))
    return feq


def mrt_tables(Q, basis, relax_mode, tau):
    M, T, om = np.zeros((Q, Q)), np.zeros((Q, Q)), np.zeros(Q)
    err = C.create_string_buffer(512)
    rc = lib().ref_mrt_tables(C.c_int(Q), C.c_int(BASES[basis]), C.c_int(RELAX_MODES[relax_mode]), C.c_double(tau),
                              _d(M), _d(T), _d(om), err, C.c_int(512))
    if rc:
        raise ValueError(err.value.decode())
    return M, T, om


def thermal_wall_point(scaling, wall_temperature, f, g):
    """One destination DoF of ThermalBounceBack<3>::calculateBoundaryValues (call sequence restated in ref_driver.cpp,
    arithmetic from the reference's headers).  f, g: (45,) arrays, modified in place; returns True if re-equilibrated."""
    return bool(lib().ref_thermal_wall_point(C.c_double(scaling), C.c_double(wall_temperature), _d(f), _d(g)))


def select_collision_range(name, scaling, f, a, b, viscosity, dt, rho_scratch, u_scratch, scheme="BGK_STANDARD",
                           equilibrium="BGK_EQUILIBRIUM"):
    """selectCollision(f) on the DoF range [a, b) of f (Q, stride), in place: what one MPI rank of the reference does on its
    owned DoFs.  rho_scratch (b-a), u_scratch (D, b-a): per-rank outputs.  Thread-safe (ctypes releases the GIL); returns
    the status code.  Used by bench.py's CPU baseline."""
    Q, stride = f.shape
    n = b - a
    fv = np.zeros(3)
    err = C.create_string_buffer(256)
    base = f.ctypes.data + 8 * a
    return lib().ref_select_collision(
        name.encode(), C.c_double(scaling), C.c_int(SCHEMES[scheme]), C.c_int(EQUILIBRIA[equilibrium]), C.c_int(0), C.c_int(0), _d(fv),
        C.c_int(0), C.c_int(0), C.c_double(viscosity), C.c_double(dt), C.c_int(0), C.c_int(0), C.c_double(1.4), C.c_int(0),
        C.c_double(1.0), C.c_int(0), C.c_int64(n), C.c_int64(stride), C.cast(base, _dp), None, _d(rho_scratch), _d(u_scratch),
        None, None, err, C.c_int(256))


# ---- the reference's own ExponentialFilter (oracle/_ref/libnatrium_ref_filter.so, oracle/ref_filter_driver.cpp) ----
_SO_FILTER = os.path.join(_HERE, "_ref", "libnatrium_ref_filter.so")
_LIB_FILTER = None


def filter_available():
    build()
    return os.path.exists(_SO_FILTER)


def _filter_lib():
    global _LIB_FILTER
    if _LIB_FILTER is None:
        if not filter_available():
            raise RuntimeError("oracle/_ref/libnatrium_ref_filter.so is missing and /root/reference is not present to build it")
        _LIB_FILTER = C.CDLL(_SO_FILTER)
    return _LIB_FILTER


def exponential_filter(dim, p, alpha, s, Nc, by_sum=False, cell_dofs=None, v=None):
    """natrium::ExponentialFilter<dim>(alpha, s, Nc, by_sum, QGaussLobatto(p+1), Lagrange element on the same nodes): returns
    (getProjectToLegendre(), getProjectFromLegendre()); with cell_dofs [n_cells, (p+1)^dim] and v, applyFilter runs in place
    on v over the cells in the given order (element-local numbering lexicographic, x fastest)."""
    n = (p + 1) ** dim
    to, fr = np.zeros((n, n)), np.zeros((n, n))
    n_cells, cd_p, v_p = 0, None, None
    if cell_dofs is not None:
        cd = np.ascontiguousarray(cell_dofs, dtype=np.int32)
        assert cd.shape[1] == n and v is not None and v.dtype == np.float64 and v.flags.c_contiguous
        n_cells, cd_p, v_p = cd.shape[0], cd.ctypes.data_as(C.POINTER(C.c_int32)), _d(v)
    rc = _filter_lib().ref_exponential_filter(C.c_int(dim), C.c_int(p), C.c_double(alpha), C.c_double(s), C.c_int(Nc), C.c_int(1 if by_sum else 0),
                                              _d(to), _d(fr), C.c_int64(n_cells), cd_p, v_p)
    if rc:
        raise ValueError((dim, p))
    return to, fr


# ---- the reference's own ThermalBounceBack (oracle/_ref/libnatrium_ref_walls.so, oracle/ref_walls_driver.cpp) ----
_SO_WALLS = os.path.join(_HERE, "_ref", "libnatrium_ref_walls.so")
_LIB_WALLS = None


def walls_available():
    build()
    return os.path.exists(_SO_WALLS)


def thermal_bounce_back(scaling, f, g, dest_index, dest_direction, wall_temperature, n=None):
    """natrium::ThermalBounceBack<3>::calculateBoundaryValues (D3Q45) for every hit of the list, in list order, in place on
    f and g ([45, stride])."""
    global _LIB_WALLS
    if _LIB_WALLS is None:
        if not walls_available():
            raise RuntimeError("oracle/_ref/libnatrium_ref_walls.so is missing and /root/reference is not present to build it")
        _LIB_WALLS = C.CDLL(_SO_WALLS)
    Q, stride = f.shape
    assert Q == 45 and g.shape == f.shape
    n = stride if n is None else n
    idx = np.ascontiguousarray(dest_index, dtype=np.int32)
    dr = np.ascontiguousarray(dest_direction, dtype=np.int32)
    rc = _LIB_WALLS.ref_thermal_bounce_back(C.c_double(scaling), C.c_int64(n), C.c_int64(stride), _d(f), _d(g), C.c_int64(len(idx)),
                                            idx.ctypes.data_as(C.POINTER(C.c_int32)), dr.ctypes.data_as(C.POINTER(C.c_int32)),
                                            C.c_double(wall_temperature))
    if rc:
        raise RuntimeError(rc)
