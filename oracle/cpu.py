"""ctypes front-end of the C oracle (oracle/natrium_oracle.c).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_dp = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)


def build(force=False):
    so = os.path.join(_HERE, "_build", "liboracle.so")
    srcs = [os.path.join(_HERE, s) for s in ("natrium_oracle.c", "entropic_oracle.c", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_num_threads.restype = C.c_int
    return _LIB


def _d(a):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(_dp)


class Stencil:
    def __init__(self, name, scaling=1.0):
        from . import stencils
        self.name = name
        self.e, self.w, self.cs2, self.max_speed = stencils.make(name, scaling)
        self.e = np.ascontiguousarray(self.e)
        self.w = np.ascontiguousarray(self.w)
        self.Q, self.D = self.e.shape
        self.scaling = float(scaling)


def spmv_csr(m, x, y=None, add=False):
    """y (+)= m @ x with the oracle's Epetra-order row loop; m is scipy CSR."""
    n = m.shape[0]
    if y is None:
        y = np.zeros(n)
    rowptr = np.ascontiguousarray(m.indptr, dtype=np.int64)
    col = np.ascontiguousarray(m.indices, dtype=np.int32)
    val = np.ascontiguousarray(m.data, dtype=np.float64)
    lib().orc_spmv_csr(C.c_int64(n), rowptr.ctypes.data_as(_i64p), col.ctypes.data_as(_i32p), _d(val),
                       _d(x), _d(y), C.c_int(1 if add else 0))
    return y


def stream(blocks, f, n_owned=None):
    """f: (Q, stride) array.  Reference order: f_tmp = copy(f); f[1:] = M f_tmp[1:]
    (CFDSolver.cpp:671-672).  Returns the new array (input untouched)."""
    Q = f.shape[0]
    f_tmp = f.copy()
    out = f.copy()
    seen = set()
    for (bi, bj) in sorted(blocks.keys()):
        m = blocks[(bi, bj)]
        n = m.shape[0]
        y = out[bi + 1, :n]
        spmv_csr(m, np.ascontiguousarray(f_tmp[bj + 1]), y, add=(bi in seen))
        seen.add(bi)
    return out


def collide_bgk(st, f, viscosity, dt, equilibrium=0, in_init=False, u_init=None, n=None):
    """In-place collideAll (f only).  Returns (rho, u[D,n], status)."""
    Q, stride = f.shape
    n = stride if n is None else n
    rho = np.zeros(n)
    u = np.zeros((st.D, n)) if u_init is None else np.ascontiguousarray(u_init, dtype=np.float64)
    rc = lib().orc_collide_bgk(C.c_int(st.D), C.c_int(Q), C.c_int64(n), C.c_int64(stride), _d(f), _d(rho), _d(u),
                               _d(st.e), _d(st.w), C.c_double(st.scaling), C.c_double(st.cs2),
                               C.c_double(viscosity), C.c_double(dt), C.c_int(equilibrium),
                               C.c_int(1 if in_init else 0))
    return rho, u, rc


def collide_bgk_fg(st, f, g, viscosity, dt, equilibrium=1, gamma=1.4, prandtl=None, sutherland=False, n=None):
    """In-place collideAll (f and g).  Returns (rho, u, T, maskShockSensor, status).
    ``prandtl=None`` mirrors isPrandtlNumberSet()==false with getPrandtlNumber()==1 default."""
    Q, stride = f.shape
    n = stride if n is None else n
    rho, T, mss = np.zeros(n), np.zeros(n), np.zeros(n)
    u = np.zeros((st.D, n))
    rc = lib().orc_collide_bgk_fg(
        C.c_int(st.D), C.c_int(Q), C.c_int64(n), C.c_int64(stride), _d(f), _d(g), _d(rho), _d(u), _d(T), _d(mss),
        _d(st.e), _d(st.w), C.c_double(st.scaling), C.c_double(st.cs2), C.c_double(viscosity), C.c_double(dt),
        C.c_int(equilibrium), C.c_double(gamma), C.c_int(0 if prandtl is None else 1),
        C.c_double(1.0 if prandtl is None else prandtl), C.c_int(1 if sutherland else 0), C.c_int(0))
    return rho, u, T, mss, rc


_GOLDEN = os.path.join(os.path.dirname(_HERE), "tests", "golden", "mrt_tables.npz")
_MRT_NAMES = {"DELLAR_D2Q9": "MRTDellarD2Q9", "LALLEMAND_D2Q9": "MRTLallemandD2Q9", "DHUMIERES_D3Q19": "MRTDHumieresD3Q19"}


def mrt_tables(basis):
    """(M, T) of AuxiliaryMRTFunctions::make_M / make_T: the reference's own literals, extracted once by
    tests/golden/make_mrt_golden.py (AuxiliaryMRTFunctions.cpp:15-205)."""
    g = np.load(_GOLDEN)
    n = _MRT_NAMES[basis]
    return np.ascontiguousarray(g[n + "_moment_trafo"]), np.ascontiguousarray(g[n + "_inverse_trafo"])


def mrt_diag(tau, basis, relax_mode="RELAX_FULL"):
    """make_diag (AuxiliaryMRTFunctions.cpp:226-404)."""
    if basis in ("DELLAR_D2Q9", "LALLEMAND_D2Q9"):
        d = np.full(9, 1.0 / tau)
        if relax_mode == "RELAX_FULL":
            d[6] = d[7] = d[8] = 1.0
        elif relax_mode == "DELLAR_RELAX_ONLY_N":
            assert basis == "DELLAR_D2Q9"
            d[8] = 1.0
        else:
            raise ValueError("MRT relaxation not defined")
        return d
    d = np.full(19, 1.0 / tau)
    if relax_mode == "RELAX_FULL":
        for i in (4, 6, 8, 10, 12, 2, 16, 17, 18):
            d[i] = 1.0
    elif relax_mode == "RELAX_DHUMIERES_PAPER":
        s9 = s13 = 1.0 / tau
        s1, s2, s10, s4, s16 = 1.19, 1.4, 1.4, 1.2, 1.98
        d = np.array([0, s1, s2, 0, s4, 0, s4, 0, s4, s9, s10, s9, s10, s13, s13, s13, s16, s16, s16], dtype=np.float64)
    else:
        raise ValueError("MRT relaxation not defined")
    return d


_ADV_SCHEMES = {"BGK_STANDARD": 0, "BGK_REGULARIZED": 1, "MRT_STANDARD": 2}
_FORCE_TYPES = {"NO_FORCING": 0, "SHIFTING_VELOCITY": 1, "EXACT_DIFFERENCE": 2, "GUO": 3}


def collide_advanced(st, f, viscosity, dt, scheme="BGK_STANDARD", equilibrium=0, in_init=False, u_init=None,
                     force=None, force_type="NO_FORCING", mrt_basis=None, relax_mode="RELAX_FULL", n=None):
    """In-place collideAll (f only) for every scheme of selectCollision on the path, with the external-force hooks.
    Returns (rho, u_scaled, status); status -2 / -3 mirror the NATriuMException / NotImplementedException of the
    force helpers."""
    Q, stride = f.shape
    n = stride if n is None else n
    rho = np.zeros(n)
    u = np.zeros((st.D, n)) if u_init is None else np.ascontiguousarray(u_init, dtype=np.float64)
    M = T = om = np.zeros(1)
    if scheme == "MRT_STANDARD":
        M, T = mrt_tables(mrt_basis)
        om = mrt_diag(viscosity / (dt * st.cs2) + 0.5, mrt_basis, relax_mode)
    fv = np.zeros(3)
    if force is not None:
        fv[:len(force)] = force
    rc = lib().orc_collide_advanced_f(
        C.c_int(st.D), C.c_int(Q), C.c_int64(n), C.c_int64(stride), _d(f), _d(rho), _d(u), _d(st.e), _d(st.w),
        C.c_double(st.scaling), C.c_double(st.cs2), C.c_double(viscosity), C.c_double(dt), C.c_int(equilibrium),
        C.c_int(_ADV_SCHEMES[scheme]), C.c_int(1 if in_init else 0), C.c_int(0 if force is None else 1),
        C.c_int(_FORCE_TYPES[force_type]), _d(fv), _d(M), _d(T), _d(om))
    return rho, u, rc


def collide_bgk_fg_forced(st, f, g, viscosity, dt, force, force_type, equilibrium=1, gamma=1.4, prandtl=None,
                          sutherland=False, n=None):
    """collide_bgk_fg with hasExternalForce() == true."""
    Q, stride = f.shape
    n = stride if n is None else n
    rho, T, mss = np.zeros(n), np.zeros(n), np.zeros(n)
    u = np.zeros((st.D, n))
    fv = np.zeros(3)
    fv[:len(force)] = force
    rc = lib().orc_collide_bgk_fg_forced(
        C.c_int(st.D), C.c_int(Q), C.c_int64(n), C.c_int64(stride), _d(f), _d(g), _d(rho), _d(u), _d(T), _d(mss),
        _d(st.e), _d(st.w), C.c_double(st.scaling), C.c_double(st.cs2), C.c_double(viscosity), C.c_double(dt),
        C.c_int(equilibrium), C.c_double(gamma), C.c_int(0 if prandtl is None else 1),
        C.c_double(1.0 if prandtl is None else prandtl), C.c_int(1 if sutherland else 0), C.c_int(0),
        C.c_int(_FORCE_TYPES[force_type]), _d(fv))
    return rho, u, T, mss, rc


def apply_wall_hits(st, f, g, dest_index, dest_dir, kind, value):
    """SemiLagrangianBoundaryHandler::apply over a flattened hit list, in place on f (post-stream) and g."""
    di = np.ascontiguousarray(dest_index, dtype=np.int32)
    dd = np.ascontiguousarray(dest_dir, dtype=np.int32)
    kk = np.ascontiguousarray(kind, dtype=np.int32)
    vv = np.ascontiguousarray(value, dtype=np.float64)
    return lib().orc_apply_wall_hits(C.c_int(st.D), C.c_int(st.Q), C.c_int64(f.shape[1]), _d(f), None if g is None else _d(g),
                                     C.c_int64(len(di)), di.ctypes.data_as(_i32p), dd.ctypes.data_as(_i32p),
                                     kk.ctypes.data_as(_i32p), _d(vv), _d(st.e), _d(st.w), C.c_double(st.scaling), C.c_double(st.cs2))


def stabilizer_matrix(name, with_e=False):
    """The reference's PseudoEntropicStabilizer literals (tests/golden/stabilizer_tables.npz,
    PseudoEntropicStabilizer.cpp:27-150)."""
    g = np.load(os.path.join(os.path.dirname(_HERE), "tests", "golden", "stabilizer_tables.npz"))
    return np.ascontiguousarray(g["d3q19" if name == "D3Q19" else ("d2q9_with_e" if with_e else "d2q9")])


def apply_stabilizer(f, A, n=None):
    """PseudoEntropicStabilizer::apply in place."""
    Q, stride = f.shape
    n = stride if n is None else n
    A = np.ascontiguousarray(A, dtype=np.float64)
    lib().orc_apply_stabilizer(C.c_int(Q), C.c_int64(n), C.c_int64(stride), _d(f), _d(A))


def collide_entropic(st, f, viscosity, dt, scheme, in_init=False, u_init=None, rho_prev=None, n=None):
    """Legacy entropic collideAll in place.  scheme: "KBC_STANDARD" (D2Q9, D3Q15; KBCStandard.cpp:88-1028) or
    "MRT_ENTROPIC" (D3Q19; MRTEntropic.cpp:167-305).  tau is the legacy nu/(dt cs2).  Returns (rho, u_scaled, status)."""
    Q, stride = f.shape
    n = stride if n is None else n
    tau = viscosity / (dt * st.cs2)
    rho = np.ones(n) if rho_prev is None else np.ascontiguousarray(rho_prev, dtype=np.float64).copy()
    u = np.zeros((st.D, n)) if u_init is None else np.ascontiguousarray(u_init, dtype=np.float64)
    L = lib()
    if scheme == "KBC_STANDARD" and (st.D, Q) == (2, 9):
        rc = L.orc_collide_kbc_d2q9(C.c_int64(n), C.c_int64(stride), _d(f), _d(rho), _d(u), C.c_double(st.scaling),
                                    C.c_double(tau), C.c_int(1 if in_init else 0))
    elif scheme == "KBC_STANDARD" and (st.D, Q) == (3, 15):
        rc = L.orc_collide_kbc_d3q15(C.c_int64(n), C.c_int64(stride), _d(f), _d(rho), _d(u), C.c_double(st.scaling),
                                     C.c_double(st.cs2), C.c_double(tau), C.c_int(1 if in_init else 0))
    elif scheme == "MRT_ENTROPIC" and (st.D, Q) == (3, 19):
        rc = L.orc_collide_mrt_entropic_d3q19(C.c_int64(n), C.c_int64(stride), _d(f), _d(rho), _d(u),
                                              C.c_double(st.scaling), C.c_double(tau), C.c_int(1 if in_init else 0))
    else:
        raise ValueError(f"{scheme} is not implemented for D{st.D}Q{Q} in the reference")
    return rho, u, rc


def mrt_entropic_tables():
    tm, invm = np.zeros((19, 19)), np.zeros((19, 19))
    lib().orc_mrt_entropic_tables(_d(tm), _d(invm))
    return tm, invm


def equilibrium(st, rho, u_unscaled, T=1.0, kind=0):
    feq = np.zeros(st.Q)
    uu = np.zeros(3)
    uu[:st.D] = u_unscaled
    lib().orc_equilibrium(C.c_int(st.D), C.c_int(st.Q), _d(st.e), _d(st.w), C.c_double(st.scaling),
                          C.c_double(st.cs2), C.c_int(kind), C.c_double(rho), _d(uu), C.c_double(T), _d(feq))
    return feq


def legacy_feq(st, rho, u):
    feq = np.zeros(st.Q)
    uu = np.ascontiguousarray(u, dtype=np.float64)
    lib().orc_legacy_feq(C.c_int(st.D), C.c_int(st.Q), _d(st.e), _d(st.w), C.c_double(st.cs2),
                         C.c_double(rho), _d(uu), _d(feq))
    return feq


def legacy_collide_single_point(st, f, tau_legacy):
    f = np.ascontiguousarray(f, dtype=np.float64).copy()
    lib().orc_legacy_collide_single_point(C.c_int(st.D), C.c_int(st.Q), _d(st.e), _d(st.w), C.c_double(st.cs2),
                                          C.c_double(tau_legacy), _d(f))
    return f


class _Blocks:
    """Keeps contiguous CSR arrays alive and exposes pointer tables for orc_step_*."""

    def __init__(self, blocks):
        keys = sorted(blocks.keys())
        self.n = len(keys)
        self.row = np.array([k[0] for k in keys], dtype=np.int32)
        self.col = np.array([k[1] for k in keys], dtype=np.int32)
        self._rp = [np.ascontiguousarray(blocks[k].indptr, dtype=np.int64) for k in keys]
        self._ci = [np.ascontiguousarray(blocks[k].indices, dtype=np.int32) for k in keys]
        self._va = [np.ascontiguousarray(blocks[k].data, dtype=np.float64) for k in keys]
        self.rowptr = (_i64p * self.n)(*[a.ctypes.data_as(_i64p) for a in self._rp])
        self.colidx = (_i32p * self.n)(*[a.ctypes.data_as(_i32p) for a in self._ci])
        self.val = (_dp * self.n)(*[a.ctypes.data_as(_dp) for a in self._va])
        self.nnz = int(sum(len(a) for a in self._va))


class ReferenceOrderStepper:
    """CFDSolver::stream()+collide() in reference order on the CPU (CFDSolver.cpp:659-843;
    CompressibleCFDSolver.h:181-314,739-788).  Used as checker and as the CPU baseline."""

    def __init__(self, st, blocks, n, viscosity, dt, equilibrium=0, with_g=False, gamma=1.4,
                 prandtl=None, sutherland=False):
        self.st, self.n, self.nu, self.dt, self.eq = st, n, viscosity, dt, equilibrium
        self.with_g, self.gamma, self.prandtl, self.sutherland = with_g, gamma, prandtl, sutherland
        self.b = _Blocks(blocks)
        self.tmp = np.zeros((st.Q, n))
        self.rho = np.zeros(n)
        self.u = np.zeros((st.D, n))
        self.T = np.zeros(n)
        self.mss = np.zeros(n)

    def step(self, f, g=None):
        st, b = self.st, self.b
        if not self.with_g:
            return lib().orc_step_f(
                C.c_int(st.D), C.c_int(st.Q), C.c_int64(self.n), C.c_int64(f.shape[1]), _d(f), _d(self.tmp),
                C.c_int(b.n), b.row.ctypes.data_as(_i32p), b.col.ctypes.data_as(_i32p), b.rowptr, b.colidx, b.val,
                _d(self.rho), _d(self.u), _d(st.e), _d(st.w), C.c_double(st.scaling), C.c_double(st.cs2),
                C.c_double(self.nu), C.c_double(self.dt), C.c_int(self.eq))
        return lib().orc_step_fg(
            C.c_int(st.D), C.c_int(st.Q), C.c_int64(self.n), C.c_int64(f.shape[1]), _d(f), _d(g), _d(self.tmp),
            C.c_int(b.n), b.row.ctypes.data_as(_i32p), b.col.ctypes.data_as(_i32p), b.rowptr, b.colidx, b.val,
            _d(self.rho), _d(self.u), _d(self.T), _d(self.mss), _d(st.e), _d(st.w), C.c_double(st.scaling),
            C.c_double(st.cs2), C.c_double(self.nu), C.c_double(self.dt), C.c_int(self.eq),
            C.c_double(self.gamma), C.c_int(0 if self.prandtl is None else 1),
            C.c_double(1.0 if self.prandtl is None else self.prandtl), C.c_int(1 if self.sutherland else 0))


def num_threads():
    return int(lib().orc_num_threads())


def set_num_threads(n):
    lib().orc_set_num_threads(C.c_int(int(n)))
