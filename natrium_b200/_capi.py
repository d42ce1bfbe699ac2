"""ctypes binding of libnatrium_b200 (include/natrium_b200.h).

The CUDA library is the product: there is no CPU or eager-PyTorch fallback.  Importing this
module never needs a GPU (so symbol/export tests run on CPU), but creating a context on a
machine without a CUDA device raises, and a missing shared library raises at import.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# NB200_LIB selects another build of the same library (tuning experiments: variants compiled with other -D flags)
LIB_PATH = os.environ.get("NB200_LIB") or os.path.join(_HERE, "libnatrium_b200.so")

NB200_OK = 0
NB200_ERR_ARG = -1
NB200_ERR_CUDA = -2
NB200_ERR_NCCL = -3
NB200_ERR_UNSUPPORTED = -4
NB200_ERR_DENSITY = -5
NB200_ERR_NO_DEVICE = -6

BGK_STANDARD, KBC_STANDARD, MRT_ENTROPIC, BGK_REGULARIZED, MRT_STANDARD = 0, 1, 2, 3, 4
NO_FORCING, SHIFTING_VELOCITY, EXACT_DIFFERENCE, GUO = 0, 1, 2, 3
WALL_VELOCITY_NEQ_BOUNCE_BACK, WALL_THERMAL_BOUNCE_BACK = 0, 1
BGK_EQUILIBRIUM, QUARTIC_EQUILIBRIUM = 0, 1
FORMAT_ELL, FORMAT_DICT, FORMAT_DICT_UNSTAGED = 0, 1, 2


class CollisionParams(C.Structure):
    _fields_ = [("scheme", C.c_int32), ("equilibrium", C.c_int32), ("with_g", C.c_int32), ("in_init", C.c_int32),
                ("viscosity", C.c_double), ("dt", C.c_double), ("gamma", C.c_double),
                ("prandtl_set", C.c_int32), ("sutherland_set", C.c_int32), ("prandtl", C.c_double),
                ("has_external_force", C.c_int32), ("force_type", C.c_int32), ("force", C.c_double * 3)]


class NatriumB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libnatrium_b200 error {code}: {msg}")
        self.code = code


class CollisionException(NatriumB200Error):
    """Mirror of natrium::CollisionException (density < 1e-10 / model not implemented)."""


_vp, _dp = C.c_void_p, C.POINTER(C.c_double)
_i32p, _i64p = C.POINTER(C.c_int32), C.POINTER(C.c_int64)

# every symbol include/natrium_b200.h declares: (restype, argtypes)
SIGNATURES = {
    "nb200_get_unique_id": (C.c_int, [_vp]),
    "nb200_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int, _vp]),
    "nb200_destroy": (None, [_vp]),
    "nb200_last_error": (C.c_char_p, [_vp]),
    "nb200_set_stencil": (C.c_int, [_vp, C.c_int, C.c_int, _dp, _dp, C.c_double, C.c_double]),
    "nb200_set_layout": (C.c_int, [_vp, C.c_int64, C.c_int64, C.c_int]),
    "nb200_set_dof_order": (C.c_int, [_vp, C.c_int64, _i32p]),
    "nb200_set_dof_grid": (C.c_int, [_vp, C.c_int, _i32p, _i32p, C.c_int]),
    "nb200_grid_info": (C.c_int, [_vp, _i64p]),
    "nb200_upload_block_csr": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int64, _i64p, _i32p, _dp]),
    "nb200_finalize_matrix": (C.c_int, [_vp]),
    "nb200_set_matrix_format": (C.c_int, [_vp, C.c_int, C.c_double]),
    "nb200_matrix_format_info": (C.c_int, [_vp, _i64p, _dp]),
    "nb200_staging_info": (C.c_int, [_vp, _i64p]),
    "nb200_set_halo": (C.c_int, [_vp, C.c_int, _i32p, _i64p, _i32p, _i64p]),
    "nb200_upload_population": (C.c_int, [_vp, C.c_int, C.c_int, _dp, C.c_int64]),
    "nb200_download_population": (C.c_int, [_vp, C.c_int, C.c_int, _dp, C.c_int64]),
    "nb200_upload_populations": (C.c_int, [_vp, C.c_int, _dp, C.c_int64]),
    "nb200_download_populations": (C.c_int, [_vp, C.c_int, _dp, C.c_int64]),
    "nb200_upload_populations_async": (C.c_int, [_vp, C.c_int, _vp, C.c_int64]),
    "nb200_download_populations_async": (C.c_int, [_vp, C.c_int, _vp, C.c_int64]),
    "nb200_upload_velocity": (C.c_int, [_vp, _dp, C.c_int64]),
    "nb200_upload_density": (C.c_int, [_vp, _dp, C.c_int64]),
    "nb200_set_collision": (C.c_int, [_vp, C.POINTER(CollisionParams)]),
    "nb200_set_mrt": (C.c_int, [_vp, C.c_int, _dp, _dp, _dp]),
    "nb200_set_post_collision_matrix": (C.c_int, [_vp, C.c_int, _dp]),
    "nb200_apply_post_collision": (C.c_int, [_vp]),
    "nb200_set_wall_hits": (C.c_int, [_vp, C.c_int64, _i32p, _i32p, _i32p, _dp]),
    "nb200_set_filter": (C.c_int, [_vp, C.c_int64, C.c_int, _i32p, _dp, _dp, _dp, C.c_int]),
    "nb200_apply_filter": (C.c_int, [_vp, C.c_int]),
    "nb200_set_iteration": (C.c_int, [_vp, C.c_int64]),
    "nb200_filter_info": (C.c_int, [_vp, _i64p]),
    "nb200_update_ghosted": (C.c_int, [_vp]),
    "nb200_stream": (C.c_int, [_vp, C.c_int]),
    "nb200_collide": (C.c_int, [_vp]),
    "nb200_step": (C.c_int, [_vp, C.c_int]),
    "nb200_step_host": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int64, C.c_int]),
    "nb200_step_host_fg": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int64]),
    "nb200_download_moments": (C.c_int, [_vp, _dp, _dp, _dp, _dp, C.c_int64]),
    "nb200_conserved": (C.c_int, [_vp, _dp]),
    "nb200_synchronize": (C.c_int, [_vp]),
    "nb200_timer_start": (C.c_int, [_vp]),
    "nb200_timer_stop": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    "nb200_kernel_launches": (C.c_int64, [_vp]),
    "nb200_matrix_info": (C.c_int, [_vp, _i64p, _i64p, _i64p]),
    "nb200_stream_handle": (_vp, [_vp]),
}

_lib = None


def load():
    """dlopen the in-tree library; fails loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(make -C natrium_b200/csrc). There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def _dptr(a):
    return a.ctypes.data_as(_dp)


def _as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Context:
    """Thin OO wrapper over one nb200_ctx.  All arrays are numpy (host) arrays."""

    def __init__(self, device=0, rank=0, nranks=1, unique_id=None):
        self.lib = load()
        self._h = _vp()
        uid = None
        if unique_id is not None:
            uid = C.create_string_buffer(bytes(unique_id), 128)
        rc = self.lib.nb200_create(C.byref(self._h), device, rank, nranks, C.cast(uid, _vp) if uid is not None else None)
        if rc != NB200_OK:
            self._h = None
            reason = {NB200_ERR_NO_DEVICE: "no CUDA device visible (there is no CPU fallback)",
                      NB200_ERR_NCCL: "NCCL initialisation failed", NB200_ERR_CUDA: "CUDA initialisation failed"}.get(rc, "bad argument")
            raise NatriumB200Error(rc, reason)
        self.D = self.Q = 0
        self.n_owned = self.n_ghost = 0
        self.with_g = False

    @staticmethod
    def unique_id():
        buf = C.create_string_buffer(128)
        rc = load().nb200_get_unique_id(C.cast(buf, _vp))
        if rc != NB200_OK:
            raise NatriumB200Error(rc, "ncclGetUniqueId failed")
        return buf.raw

    def _check(self, rc):
        if rc == NB200_OK:
            return
        msg = self.lib.nb200_last_error(self._h).decode()
        if rc in (NB200_ERR_DENSITY, NB200_ERR_UNSUPPORTED):
            raise CollisionException(rc, msg)
        raise NatriumB200Error(rc, msg)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.nb200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- static data
    def set_stencil(self, e_scaled, w, scaling, cs2_scaled):
        e = _as_f64(e_scaled)
        w = _as_f64(w)
        self.Q, self.D = e.shape
        self._check(self.lib.nb200_set_stencil(self._h, self.D, self.Q, _dptr(e), _dptr(w), scaling, cs2_scaled))

    def set_layout(self, n_owned, n_ghost=0, with_g=False):
        self._check(self.lib.nb200_set_layout(self._h, n_owned, n_ghost, 1 if with_g else 0))
        self.n_owned, self.n_ghost, self.with_g = int(n_owned), int(n_ghost), bool(with_g)

    def set_dof_order(self, order):
        order = np.ascontiguousarray(order, dtype=np.int32)
        self._check(self.lib.nb200_set_dof_order(self._h, len(order), order.ctypes.data_as(_i32p)))

    def set_dof_grid(self, dims, coords, fe_order=0):
        """Structure hint: integer grid coordinates [(n_owned + n_ghost), dim] of every local DoF (nb200_set_dof_grid)."""
        dims = np.ascontiguousarray(dims, dtype=np.int32)
        coords = np.ascontiguousarray(coords, dtype=np.int32)
        assert coords.ndim == 2 and coords.shape[1] == len(dims)
        self._check(load().nb200_set_dof_grid(self._h, C.c_int(len(dims)), dims.ctypes.data_as(_i32p),
                                              coords.ctypes.data_as(_i32p), C.c_int(int(fe_order))))

    def grid_info(self):
        out = np.zeros(8, dtype=np.int64)
        self._check(load().nb200_grid_info(self._h, out.ctypes.data_as(_i64p)))
        keys = ("in_use", "tiles", "box_rows", "generic_rows", "boxes", "passes", "pass_capacity", "grid_points")
        return {k: int(v) for k, v in zip(keys, out)}

    def upload_block_csr(self, bi, bj, rowptr, col, val):
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
        col = np.ascontiguousarray(col, dtype=np.int32)
        val = _as_f64(val)
        self._check(self.lib.nb200_upload_block_csr(self._h, bi, bj, len(rowptr) - 1, rowptr.ctypes.data_as(_i64p),
                                                    col.ctypes.data_as(_i32p), _dptr(val)))

    def finalize_matrix(self):
        self._check(self.lib.nb200_finalize_matrix(self._h))

    def set_matrix_format(self, fmt=FORMAT_DICT, value_dedup_tol=1e-14):
        self._check(self.lib.nb200_set_matrix_format(self._h, fmt, value_dedup_tol))

    def matrix_format_info(self):
        out = (C.c_int64 * 6)()
        tol = C.c_double()
        self._check(self.lib.nb200_matrix_format_info(self._h, out, C.byref(tol)))
        st = (C.c_int64 * 5)()
        self._check(self.lib.nb200_staging_info(self._h, st))
        return dict(format="dict" if out[0] == FORMAT_DICT else "ell", patterns=out[1], lists=out[2], pool_bytes=out[3],
                    descriptor_bytes=out[4], classes=out[5], value_dedup_tol=tol.value, staged=bool(st[0]),
                    staged_values=st[1], staging_passes=st[2], largest_pass=st[3], pass_capacity=st[4])

    def set_halo(self, nbr_rank, send_off, send_idx, recv_off):
        nbr = np.ascontiguousarray(nbr_rank, dtype=np.int32)
        so = np.ascontiguousarray(send_off, dtype=np.int64)
        si = np.ascontiguousarray(send_idx, dtype=np.int32)
        ro = np.ascontiguousarray(recv_off, dtype=np.int64)
        self._check(self.lib.nb200_set_halo(self._h, len(nbr), nbr.ctypes.data_as(_i32p), so.ctypes.data_as(_i64p),
                                            si.ctypes.data_as(_i32p), ro.ctypes.data_as(_i64p)))

    # ---- populations
    def upload_population(self, which, q, host):
        host = _as_f64(host)
        self._check(self.lib.nb200_upload_population(self._h, which, q, _dptr(host), host.shape[0]))

    def download_population(self, which, q):
        out = np.empty(self.n_owned)
        self._check(self.lib.nb200_download_population(self._h, which, q, _dptr(out), self.n_owned))
        return out

    def upload_populations(self, which, host):
        host = _as_f64(host)
        assert host.shape == (self.Q, self.n_owned), host.shape
        self._check(self.lib.nb200_upload_populations(self._h, which, _dptr(host), self.n_owned))

    def download_populations(self, which, out=None):
        if out is None:
            out = np.empty((self.Q, self.n_owned))
        self._check(self.lib.nb200_download_populations(self._h, which, _dptr(out), self.n_owned))
        return out

    def upload_populations_async(self, which, host_ptr):
        self._check(self.lib.nb200_upload_populations_async(self._h, which, host_ptr, self.n_owned))

    def download_populations_async(self, which, host_ptr):
        self._check(self.lib.nb200_download_populations_async(self._h, which, host_ptr, self.n_owned))

    def upload_velocity(self, u):
        u = _as_f64(u)
        self._check(self.lib.nb200_upload_velocity(self._h, _dptr(u), u.shape[1]))

    def upload_density(self, rho):
        rho = _as_f64(rho)
        self._check(self.lib.nb200_upload_density(self._h, _dptr(rho), rho.shape[0]))

    # ---- operators
    def set_collision(self, viscosity, dt, scheme=BGK_STANDARD, equilibrium=BGK_EQUILIBRIUM, with_g=False,
                      in_init=False, gamma=1.4, prandtl=None, sutherland=False, force=None, force_type=NO_FORCING):
        """force: external force vector (problemDescription.getExternalForce()->getForce()) or None for
        hasExternalForce() == false; force_type: configuration.getForcingScheme()."""
        fv = (C.c_double * 3)(0.0, 0.0, 0.0)
        if force is not None:
            for j, v in enumerate(force):
                fv[j] = float(v)
        p = CollisionParams(scheme, equilibrium, 1 if with_g else 0, 1 if in_init else 0, viscosity, dt, gamma,
                            0 if prandtl is None else 1, 1 if sutherland else 0, 1.0 if prandtl is None else prandtl,
                            0 if force is None else 1, force_type, fv)
        self._check(self.lib.nb200_set_collision(self._h, C.byref(p)))

    def set_wall_hits(self, dest_index, dest_direction, kind, value):
        """Flattened HitList of the semi-Lagrangian boundary handler (nb200_set_wall_hits)."""
        di = np.ascontiguousarray(dest_index, dtype=np.int32)
        dd = np.ascontiguousarray(dest_direction, dtype=np.int32)
        kk = np.ascontiguousarray(kind, dtype=np.int32)
        vv = _as_f64(value)
        self._check(self.lib.nb200_set_wall_hits(self._h, len(di), di.ctypes.data_as(_i32p), dd.ctypes.data_as(_i32p),
                                                 kk.ctypes.data_as(_i32p), _dptr(vv)))

    def set_post_collision_matrix(self, A):
        """PseudoEntropicStabilizer matrix (Q x Q); None removes it."""
        if A is None:
            self._check(self.lib.nb200_set_post_collision_matrix(self._h, 0, None))
            return
        A = _as_f64(A)
        self._check(self.lib.nb200_set_post_collision_matrix(self._h, A.shape[0], _dptr(A)))

    def apply_post_collision(self):
        self._check(self.lib.nb200_apply_post_collision(self._h))

    def set_filter(self, cell_dofs, to_legendre, from_legendre, sigma, interval=1):
        """ExponentialFilter tables (nb200_set_filter): cell_dofs [n_cells, n] in the host's cell order; None removes it."""
        if cell_dofs is None:
            self._check(self.lib.nb200_set_filter(self._h, 0, 0, None, None, None, None, 0))
            return
        cd = np.ascontiguousarray(cell_dofs, dtype=np.int32)
        to, fr, sg = _as_f64(to_legendre), _as_f64(from_legendre), _as_f64(sigma)
        n = cd.shape[1]
        assert to.shape == (n, n) and fr.shape == (n, n) and sg.shape == (n,)
        self._check(self.lib.nb200_set_filter(self._h, cd.shape[0], n, cd.ctypes.data_as(_i32p), _dptr(to), _dptr(fr), _dptr(sg), int(interval)))

    def apply_filter(self, which=0):
        self._check(self.lib.nb200_apply_filter(self._h, which))

    def set_iteration(self, i):
        self._check(self.lib.nb200_set_iteration(self._h, int(i)))

    def filter_info(self):
        out = np.zeros(5, dtype=np.int64)
        self._check(self.lib.nb200_filter_info(self._h, out.ctypes.data_as(_i64p)))
        return dict(zip(("cells", "dofs_per_cell", "levels", "interval", "iteration"), (int(v) for v in out)))

    def set_mrt(self, M, T, omega):
        """Tables of MultipleRelaxationTime::SpecificCollisionData (make_M / make_T / make_diag)."""
        M, T, omega = _as_f64(M), _as_f64(T), _as_f64(omega)
        self._check(self.lib.nb200_set_mrt(self._h, int(len(omega)), _dptr(M), _dptr(T), _dptr(omega)))

    def update_ghosted(self):
        self._check(self.lib.nb200_update_ghosted(self._h))

    def stream(self, which=0):
        self._check(self.lib.nb200_stream(self._h, which))

    def collide(self):
        self._check(self.lib.nb200_collide(self._h))

    def step(self, n_steps=1):
        self._check(self.lib.nb200_step(self._h, n_steps))

    def step_host(self, f_in_ptr, f_out_ptr, rho_ptr, u_ptr, n_chunks=16):
        """One step with host buffers given as raw addresses (page-locked for true overlap); fence with synchronize()."""
        self._check(self.lib.nb200_step_host(self._h, f_in_ptr, f_out_ptr, rho_ptr, u_ptr, self.n_owned, n_chunks))

    def step_host_fg(self, f_in_ptr, g_in_ptr, f_out_ptr, g_out_ptr, rho_ptr, u_ptr, T_ptr):
        self._check(self.lib.nb200_step_host_fg(self._h, f_in_ptr, g_in_ptr, f_out_ptr, g_out_ptr, rho_ptr, u_ptr, T_ptr, self.n_owned))

    def synchronize(self):
        self._check(self.lib.nb200_synchronize(self._h))

    # ---- results
    def download_moments(self, want_T=False):
        n = self.n_owned
        rho, u = np.empty(n), np.empty((self.D, n))
        T = np.empty(n) if want_T else None
        s = np.empty(n) if want_T else None
        self._check(self.lib.nb200_download_moments(self._h, _dptr(rho), _dptr(u), _dptr(T) if want_T else None,
                                                    _dptr(s) if want_T else None, n))
        return (rho, u, T, s) if want_T else (rho, u)

    def conserved(self):
        out = np.zeros(5)
        self._check(self.lib.nb200_conserved(self._h, _dptr(out)))
        return out

    def timer_start(self):
        self._check(self.lib.nb200_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_float()
        self._check(self.lib.nb200_timer_stop(self._h, C.byref(ms)))
        return float(ms.value)

    def kernel_launches(self):
        return int(self.lib.nb200_kernel_launches(self._h))

    def matrix_info(self):
        nnz, nbytes, padded = C.c_int64(), C.c_int64(), C.c_int64()
        self._check(self.lib.nb200_matrix_info(self._h, C.byref(nnz), C.byref(nbytes), C.byref(padded)))
        return dict(nnz=nnz.value, device_bytes=nbytes.value, padded_entries=padded.value)
