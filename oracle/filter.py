"""CPU restatement of the reference's ExponentialFilter set-up (smoothing/ExponentialFilter.cpp).  Test infrastructure only.

The reference builds, once, on the host (deal.II):
  * projection matrices nodal FE basis <-> tensor Legendre modes from quadrature sums
    (`makeProjectionMatrices`, ExponentialFilter.cpp:30-68): quad_source(k,i) = sum_q w_q phi_i(x_q) psi_k(x_q),
    quad_legendre(k,i) = sum_q w_q psi_i(x_q) psi_k(x_q), to_legendre = quad_legendre^-1 quad_source,
    from_legendre = to_legendre^-1, with the operator's own QGaussLobatto(p+1) quadrature and its Lagrange element on the
    same Gauss-Lobatto nodes (advection/AdvectionOperator.cpp:34-39);
  * the degree of every mode (`makeDegreeVectors`, :96-137) and, inside applyFilter (:171-183), the damping factor
    sigma = exp(-alpha ((degree + 1 - Nc) / (max_degree + 1 - Nc))^s) of the modes with degree >= Nc.
The element-local numbering here is lexicographic (x fastest; FE_DGQArbitraryNodes).  A continuous FE_Q numbers its shape
functions hierarchically: the filter is invariant under any renumbering applied to the matrix columns/rows and to
cell->get_dof_indices alike, so the host hands over whatever numbering it has.

Bug-for-bug: the 3-D degree vector of the reference computes iy as `(i % (p+1) * (p+1)) / (p+1)`, which by operator
precedence is i % (p+1) = iz (:123).  Results identical to the reference's need the same degrees, so it is restated as is
(`reference_quirk=True`); `False` gives the intended (ix, iy, iz).

Pinned to the reference's own code: tests/test_oracle_vs_ref.py runs L/smoothing/ExponentialFilter.cpp, compiled from
/root/reference against oracle/ref_stubs_filter (oracle/_ref/libnatrium_ref_filter.so), beside these functions.

The Legendre basis is dealii::Polynomials::Legendre: orthonormal on [0, 1], L_k(x) = sqrt(2k+1) P_k(2x - 1); the mode index
decomposes with x slowest (evaluateLegendreND, :70-94).  The filter itself does not depend on the normalisation.
"""
import numpy as np


def gauss_lobatto_01(npts):
    """Nodes and weights of the (npts)-point Gauss-Lobatto rule on [0, 1] (dealii::QGaussLobatto<1>)."""
    n = npts - 1
    if n == 0:
        return np.array([0.5]), np.array([1.0])
    Pn = np.polynomial.legendre.Legendre.basis(n)
    x = np.concatenate([[-1.0], np.sort(np.real(Pn.deriv().roots())), [1.0]])
    for _ in range(3):       # Newton polish of the interior nodes: roots of P_n'
        xi = x[1:-1]
        d1, d2 = Pn.deriv()(xi), Pn.deriv(2)(xi)
        x[1:-1] = xi - d1 / d2
    w = 2.0 / (n * (n + 1) * Pn(x) ** 2)
    return 0.5 * (x + 1.0), 0.5 * w


def lagrange_1d(nodes, xi):
    """phi_j(xi) for all j: Lagrange polynomials on `nodes`."""
    out = np.ones(len(nodes))
    for j in range(len(nodes)):
        for m in range(len(nodes)):
            if m != j:
                out[j] *= (xi - nodes[m]) / (nodes[j] - nodes[m])
    return out


def legendre_01(k, x):
    """dealii::Polynomials::Legendre(k).value(x): orthonormal on [0, 1]."""
    return np.sqrt(2.0 * k + 1.0) * np.polynomial.legendre.Legendre.basis(k)(2.0 * np.asarray(x) - 1.0)


def _mode_indices(i, p, dim):
    """(ix, iy, iz) of Legendre mode i as evaluateLegendreND decomposes it (x slowest)."""
    n1 = p + 1
    if dim == 1:
        return (i,)
    if dim == 2:
        return (i // n1, i % n1)
    return (i // (n1 * n1), (i % (n1 * n1)) // n1, i % n1)


def projection_matrices(p, dim):
    """(to_legendre, from_legendre), [n, n] with n = (p+1)^dim, element-local numbering lexicographic (x fastest)."""
    n1 = p + 1
    n = n1 ** dim
    x1, w1 = gauss_lobatto_01(n1)
    # quadrature points, x fastest (dealii::QGaussLobatto<dim> is the tensor product with the first coordinate fastest)
    idx = np.stack(np.meshgrid(*[np.arange(n1)] * dim, indexing="ij"), axis=-1).reshape(-1, dim)[:, ::-1]   # [q, axis], x fastest
    wq = np.prod(w1[idx], axis=1)
    phi1 = np.array([lagrange_1d(x1, xq) for xq in x1])            # [point, j]
    leg1 = np.array([[legendre_01(k, xq) for k in range(n1)] for xq in x1])   # [point, k]
    phi = np.ones((n, n))       # [q, i]
    psi = np.ones((n, n))       # [q, mode]
    for i in range(n):
        ii = idx[i]             # shape function i sits on node ii (lexicographic)
        mi = _mode_indices(i, p, dim)
        for d in range(dim):
            phi[:, i] *= phi1[idx[:, d], ii[d]]
            psi[:, i] *= leg1[idx[:, d], mi[d]]
    quad_source = np.zeros((n, n))
    quad_legendre = np.zeros((n, n))
    for q in range(n):          # the reference's accumulation order: q innermost per (k, i)
        quad_source += wq[q] * np.outer(psi[q], phi[q])
        quad_legendre += wq[q] * np.outer(psi[q], psi[q])
    to_legendre = np.linalg.inv(quad_legendre) @ quad_source
    from_legendre = np.linalg.inv(to_legendre)
    return np.ascontiguousarray(to_legendre), np.ascontiguousarray(from_legendre)


def degree_vectors(p, dim, reference_quirk=True):
    n1 = p + 1
    n = n1 ** dim
    dmax, dsum = np.zeros(n, dtype=np.int64), np.zeros(n, dtype=np.int64)
    for i in range(n):
        if dim == 3 and reference_quirk:
            ix = i // (n1 * n1)
            iy = ((i % n1) * n1) // n1          # ExponentialFilter.cpp:123 as written: = i % (p+1)
            iz = i % n1
            m = (ix, iy, iz)
        else:
            m = _mode_indices(i, p, dim)
        dmax[i], dsum[i] = max(m), sum(m)
    return dmax, dsum


def damping(p, dim, alpha, s, Nc, by_sum=False, reference_quirk=True):
    """(sigma[n], damped[n]): factor of every Legendre mode and which modes applyFilter multiplies (degree >= Nc)."""
    dmax, dsum = degree_vectors(p, dim, reference_quirk)
    deg = dsum if by_sum else dmax
    max_degree = dim * p if by_sum else p
    sigma = np.ones(len(deg))
    damped = deg >= Nc
    sigma[damped] = np.exp(-alpha * ((deg[damped] + 1.0 - Nc) / (max_degree + 1.0 - Nc)) ** s)
    return sigma, damped.astype(np.uint8)


def apply_filter(cell_dofs, to_legendre, from_legendre, sigma, damped, v):
    """In place on v (1-d, all local DoFs): the C oracle's cell-by-cell loop (orc_exponential_filter)."""
    import ctypes as C
    from . import cpu
    cell_dofs = np.ascontiguousarray(cell_dofs, dtype=np.int32)
    n_cells, n = cell_dofs.shape
    assert v.dtype == np.float64 and v.flags.c_contiguous and n <= 1024
    cpu.lib().orc_exponential_filter(C.c_int64(n_cells), C.c_int(n), cell_dofs.ctypes.data_as(C.POINTER(C.c_int32)),
                                     cpu._d(np.ascontiguousarray(to_legendre)), cpu._d(np.ascontiguousarray(from_legendre)),
                                     cpu._d(np.ascontiguousarray(sigma, dtype=np.float64)),
                                     np.ascontiguousarray(damped, dtype=np.uint8).ctypes.data_as(C.POINTER(C.c_ubyte)), cpu._d(v))
    return v
