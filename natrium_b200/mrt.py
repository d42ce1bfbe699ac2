"""Host-side mirror of AuxiliaryMRTFunctions (L/collision_advanced/AuxiliaryMRTFunctions.{h,cpp}): the moment
transforms M, their inverses T and the relaxation diagonals that MultipleRelaxationTime::SpecificCollisionData
(L/collision_advanced/CollisionSchemes.h:209-236) builds from make_M / make_T / make_diag and that the host hands to
nb200_set_mrt().  In a NATriuM build the reference's own tables are passed; here the bases are rebuilt from their
published definitions in the reference's direction order (tests pin them to the reference's literals,
tests/golden/mrt_tables.npz).

  DELLAR_D2Q9      weighted Hermite basis with ghost modes N, J (Dellar 2003)       AuxiliaryMRTFunctions.cpp:15-45
  LALLEMAND_D2Q9   rho, jx, jy, pxx, pxy, e, qx, qy, eps (Lallemand & Luo 2000)     :51-80
  DHUMIERES_D3Q19  rho, e, eps, jx, qx, jy, qy, jz, qz, 3pxx, 3pixx, pww, piww, pxy, pyz, pxz, mx, my, mz
                   (d'Humieres et al. 2002)                                          :86-205
"""
import numpy as np

# RelaxMode / MomentBasis (L/utilities/ConfigNames.h:20-33)
RELAX_FULL, DELLAR_RELAX_ONLY_N, RELAX_DHUMIERES_PAPER = 0, 1, 2
DELLAR_D2Q9, LALLEMAND_D2Q9, DHUMIERES_D3Q19 = 0, 1, 2

_E_D2Q9 = np.array([[0, 0], [1, 0], [0, 1], [-1, 0], [0, -1], [1, 1], [-1, 1], [-1, -1], [1, -1]], dtype=np.float64)
_W_D2Q9 = np.array([4 / 9] + [1 / 9] * 4 + [1 / 36] * 4)
_E_D3Q19 = np.array([[0, 0, 0], [1, 0, 0], [0, 0, 1], [-1, 0, 0], [0, 0, -1], [0, -1, 0], [0, 1, 0], [1, 0, 1], [-1, 0, 1],
                     [-1, 0, -1], [1, 0, -1], [1, -1, 0], [1, 1, 0], [-1, 1, 0], [-1, -1, 0], [0, -1, 1], [0, 1, 1],
                     [0, 1, -1], [0, -1, -1]], dtype=np.float64)


def _basis_rows(basis):
    if basis == DELLAR_D2Q9:
        x, y = _E_D2Q9.T
        g = np.where(x * x + y * y == 0, 1.0, np.where(x * x + y * y == 1, -2.0, 4.0))     # Dellar's ghost vector
        return np.array([np.ones(9), x, y, 4.5 * x * x - 1.5, 9.0 * x * y, 4.5 * y * y - 1.5, g, g * x, g * y]), _W_D2Q9
    if basis == LALLEMAND_D2Q9:
        x, y = _E_D2Q9.T
        c2 = x * x + y * y
        return np.array([np.ones(9), x, y, x * x - y * y, x * y, -4 + 3 * c2, (-5 + 3 * c2) * x, (-5 + 3 * c2) * y,
                         4 - 10.5 * c2 + 4.5 * c2 * c2]), None
    if basis == DHUMIERES_D3Q19:
        x, y, z = _E_D3Q19.T
        c2 = x * x + y * y + z * z
        return np.array([np.ones(19), 19 * c2 - 30, (21 * c2 * c2 - 53 * c2 + 24) / 2, x, (5 * c2 - 9) * x, y, (5 * c2 - 9) * y,
                         z, (5 * c2 - 9) * z, 3 * x * x - c2, (3 * c2 - 5) * (3 * x * x - c2), y * y - z * z,
                         (3 * c2 - 5) * (y * y - z * z), x * y, y * z, x * z, (y * y - z * z) * x, (z * z - x * x) * y,
                         (x * x - y * y) * z]), None
    raise ValueError("MRT basis not defined")


def make_M(basis):
    """Moment transform (make_M, AuxiliaryMRTFunctions.cpp:156-187)."""
    return np.ascontiguousarray(_basis_rows(basis)[0] + 0.0)


def make_T(basis):
    """Inverse transform (make_T, :191-222).  The rows of M are orthogonal (Dellar: with respect to the lattice
    weights), so T[:, p] = w * M[p] / sum(w * M[p]^2) -- rational entries, as in the reference's literals."""
    M, w = _basis_rows(basis)
    w = np.ones(M.shape[1]) if w is None else w
    norm = np.sum(w[None, :] * M * M, axis=1)
    return np.ascontiguousarray((w[None, :] * M / norm[:, None]).T)


def make_diag(tau, basis, relax_mode=RELAX_FULL):
    """Relaxation rates (make_diag, :226-404)."""
    if basis in (DELLAR_D2Q9, LALLEMAND_D2Q9):
        d = np.full(9, 1.0 / tau)
        if relax_mode == RELAX_FULL:
            d[6:9] = 1.0
        elif relax_mode == DELLAR_RELAX_ONLY_N:
            if basis != DELLAR_D2Q9:
                raise ValueError("DELLAR_RELAX_ONLY_N is only supported for the DELLAR_D2Q9 moment basis")
            d[8] = 1.0        # the reference relaxes entry 8 (AuxiliaryMRTFunctions.cpp:270-274)
        else:
            raise ValueError("MRT relaxation not defined")
        return d
    if basis == DHUMIERES_D3Q19:
        d = np.full(19, 1.0 / tau)
        if relax_mode == RELAX_FULL:
            d[[2, 4, 6, 8, 10, 12, 16, 17, 18]] = 1.0
        elif relax_mode == RELAX_DHUMIERES_PAPER:
            s9 = s13 = 1.0 / tau
            s1, s2, s10, s4, s16 = 1.19, 1.4, 1.4, 1.2, 1.98
            d = np.array([0, s1, s2, 0, s4, 0, s4, 0, s4, s9, s10, s9, s10, s13, s13, s13, s16, s16, s16], dtype=np.float64)
        else:
            raise ValueError("MRT relaxation not defined")
        return d
    raise ValueError("MRT basis not defined")


def make_stabilizer(stencil, with_e=False):
    """Matrix of PseudoEntropicStabilizer::apply (L/dataprocessors/PseudoEntropicStabilizer.cpp:27-150): in moment space
    the conserved and second-order moments are kept and every higher moment is replaced by the linear part of its
    equilibrium closure, A = T C M.
      D2Q9  (Lallemand basis): qx = -jx, qy = -jy, eps = -(rho + e); with_e: e = -2 rho, eps = rho
      D3Q19 (d'Humieres basis): eps = -(7 rho + 11 e)/38 (w_eps = 3, w_epsj = -11/2), q = -2/3 j, pi_xx = -1/2 (3 p_xx),
            pi_ww = -1/2 p_ww, m = 0
    Tests pin the result to the reference's literals (tests/golden/stabilizer_tables.npz)."""
    if stencil in ("D2Q9", "Stencil_D2Q9"):
        M, T = make_M(LALLEMAND_D2Q9), make_T(LALLEMAND_D2Q9)
        C = np.eye(9)
        C[6], C[7], C[8] = -C[1], -C[2], 0.0
        if with_e:
            C[5] = -2.0 * np.eye(9)[0]
            C[8] = np.eye(9)[0]
        else:
            C[8] = -np.eye(9)[0] - np.eye(9)[5]
        return np.ascontiguousarray(T @ C @ M)
    if stencil in ("D3Q19", "Stencil_D3Q19"):
        if with_e:
            raise ValueError("the with-e variant exists for D2Q9 only")
        M, T = make_M(DHUMIERES_D3Q19), make_T(DHUMIERES_D3Q19)
        I = np.eye(19)
        C = I.copy()
        C[2] = -(7.0 * I[0] + 11.0 * I[1]) / 38.0
        C[4], C[6], C[8] = -2.0 / 3.0 * I[3], -2.0 / 3.0 * I[5], -2.0 / 3.0 * I[7]
        C[10], C[12] = -0.5 * I[9], -0.5 * I[11]
        C[16] = C[17] = C[18] = 0.0
        return np.ascontiguousarray(T @ C @ M)
    raise ValueError("PseudoEntropicStabilizer is only defined for D2Q9 and D3Q19")
