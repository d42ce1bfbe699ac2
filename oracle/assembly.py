"""Oracle restatement of the semi-Lagrangian streaming-matrix assembly (test infrastructure only).

Restates, for Cartesian (optionally stretched) hyper-rectangle meshes with periodic or
bounce-back faces, what the reference does once on the host:

  L/advection/SemiLagrangian.cpp:145-522   fillSparseObject (main loop, path tracking, add())
  L/advection/SemiLagrangian.cpp:525-699   faceCrossedFirst (unit-cell ray / face test, 1e-10 snapping)
  L/advection/SemiLagrangianTools.cpp:15-64 shapeFunctionValue (snap unit point, FE_Q shape values)
  L/advection/AdvectionTools.cpp:48-95     FE_Q on Gauss-Lobatto support points
  L/boundaries/PeriodicBoundary.cpp:184-   coordinatesAcrossPeriodicBoundary (pure translation)
  L/utilities/CFDSolverUtilities.cpp:92-100 dt = CFL*dx_min/(|e|_max*p^2)

deal.II pieces that are not in /root/reference are restated from their published
definition: MappingCartesian (affine per axis), FE_Q(QGaussLobatto<1>(p+1)) = tensor product
of 1D Lagrange polynomials on the GLL points of [0,1], lexicographic in (x fastest).

The output is the (Q-1)x(Q-1) block matrix as scipy CSR blocks with sorted column
indices (Epetra local order after FillComplete).  Pure-Python loops: small cases only.
"""
import math
from collections import deque

import numpy as np
import scipy.sparse as sp


def gll_nodes(p):
    """Gauss-Lobatto points on [0,1] (dealii::QGaussLobatto<1>(p+1)): endpoints + roots of P'_p."""
    n = p + 1
    if n == 2:
        return np.array([0.0, 1.0])
    # interior nodes: roots of derivative of Legendre P_p on [-1,1], Newton from Chebyshev-Lobatto guess
    x = -np.cos(np.pi * np.arange(n) / p)
    for _ in range(100):
        # Legendre P_p and P_{p-1} by recurrence
        P0 = np.ones_like(x)
        P1 = x.copy()
        for k in range(2, p + 1):
            P0, P1 = P1, ((2 * k - 1) * x * P1 - (k - 1) * P0) / k
        # P1 = P_p, P0 = P_{p-1};  (1-x^2) P_p' = p (P_{p-1} - x P_p)
        f = p * (P0 - x * P1)              # = (1-x^2) P_p'(x): zero at all GLL nodes incl. endpoints
        df = -p * (p + 1) * P1             # d/dx[(1-x^2)P_p'] = -p(p+1) P_p
        dx = f / df
        x = x - dx
        if np.max(np.abs(dx)) < 1e-16:
            break
    x[0], x[-1] = -1.0, 1.0
    x = 0.5 * (x + 1.0)
    # symmetrise
    x = 0.5 * (x + (1.0 - x[::-1]))
    return x


def lagrange_values(nodes, xi):
    """l_j(xi) = prod_{m != j} (xi - x_m) / (x_j - x_m), j = 0..p."""
    n = len(nodes)
    out = np.empty(n)
    for j in range(n):
        v = 1.0
        for m in range(n):
            if m != j:
                v *= (xi - nodes[m]) / (nodes[j] - nodes[m])
        out[j] = v
    return out


class CartesianMesh:
    """Axis-aligned hex/quad mesh given by per-axis vertex coordinates.

    ``boundary[d]`` is 'periodic' or 'wall' (VelocityNeqBounceBack-style: path is reflected,
    direction reversed, SemiLagrangian.cpp:358-384).
    """

    def __init__(self, verts, boundary=None):
        self.verts = [np.asarray(v, dtype=np.float64) for v in verts]
        self.dim = len(self.verts)
        self.n = [len(v) - 1 for v in self.verts]
        self.boundary = boundary or ["periodic"] * self.dim

    @staticmethod
    def uniform(dim, n, L=2 * math.pi, origin=0.0):
        n = [n] * dim if np.isscalar(n) else list(n)
        L = [L] * dim if np.isscalar(L) else list(L)
        return CartesianMesh([origin + L[d] * np.arange(n[d] + 1) / n[d] for d in range(dim)])

    def min_vertex_distance(self):
        """CFDSolverUtilities::getMinimumVertexDistance (CFDSolverUtilities.cpp:69-85)."""
        return min(float(np.min(np.diff(v))) for v in self.verts)

    def cells(self):
        """Cell multi-indices in lexicographic order (x fastest)."""
        rng = [range(k) for k in self.n]
        if self.dim == 2:
            return [(cx, cy) for cy in rng[1] for cx in rng[0]]
        return [(cx, cy, cz) for cz in rng[2] for cy in rng[1] for cx in rng[0]]


def calculate_timestep(mesh, p, max_speed, cfl):
    """CFDSolverUtilities::calculateTimestep (CFDSolverUtilities.cpp:92-100)."""
    return cfl * mesh.min_vertex_distance() / (max_speed * p * p)


class DofMap:
    """Continuous FE_Q(p) DoFs on the Cartesian mesh; periodic faces are NOT identified
    (periodicity is handled by path tracking, SemiLagrangian.cpp:336-352)."""

    def __init__(self, mesh, p, numbering=None):
        self.mesh, self.p = mesh, p
        self.nd = [k * p + 1 for k in mesh.n]
        self.N = int(np.prod(self.nd))
        self.numbering = None if numbering is None else np.asarray(numbering)

    def lex(self, g):
        i = 0
        for d in reversed(range(self.mesh.dim)):
            i = i * self.nd[d] + g[d]
        return i

    def index(self, g):
        i = self.lex(g)
        return int(self.numbering[i]) if self.numbering is not None else i

    def cell_dofs(self, cell):
        """global indices of the (p+1)^dim local DoFs, lexicographic (x fastest)."""
        p, dim = self.p, self.mesh.dim
        out = []
        if dim == 2:
            for b in range(p + 1):
                for a in range(p + 1):
                    out.append(self.index((cell[0] * p + a, cell[1] * p + b)))
        else:
            for c in range(p + 1):
                for b in range(p + 1):
                    for a in range(p + 1):
                        out.append(self.index((cell[0] * p + a, cell[1] * p + b, cell[2] * p + c)))
        return out

    def support_points(self):
        """(N, dim) coordinates indexed by global DoF index."""
        nodes = gll_nodes(self.p)
        axes = []
        for d in range(self.mesh.dim):
            v = self.mesh.verts[d]
            ax = np.empty(self.nd[d])
            for c in range(self.mesh.n[d]):
                ax[c * self.p:(c + 1) * self.p + 1] = v[c] + nodes * (v[c + 1] - v[c])
            axes.append(ax)
        grids = np.meshgrid(*axes, indexing="ij")
        pts_lex = np.stack([g.transpose().reshape(-1) for g in grids], axis=1)  # x fastest
        if self.numbering is None:
            return pts_lex
        out = np.empty_like(pts_lex)
        out[self.numbering] = pts_lex
        return out


def _snap(v):
    # "eliminate round-off-errors" SemiLagrangian.cpp:557-570 / SemiLagrangianTools.cpp:41-47
    if abs(v) < 1e-10:
        return 0.0
    if abs(v - 1) < 1e-10:
        return 1.0
    return v


def _face_crossed_first(mesh, cell, p_in, p_out):
    """SemiLagrangian.cpp:525-699 for a Cartesian cell.  Returns (face_id, p_boundary)."""
    face, pb, _ = face_crossed_first(mesh, cell, p_in, p_out)
    return face, pb


def face_crossed_first(mesh, cell, p_in, p_out):
    """SemiLagrangian<dim>::faceCrossedFirst (SemiLagrangian.cpp:525-699) for a Cartesian cell: the face the segment
    p_in -> p_out leaves the cell through (deal.II numbering 2*d + side, -1: p_out inside), the boundary point and
    lambda (fraction of the segment up to the face).  Returns (face_id, p_boundary, lambda)."""
    dim = mesh.dim
    x0 = [mesh.verts[d][cell[d]] for d in range(dim)]
    h = [mesh.verts[d][cell[d] + 1] - mesh.verts[d][cell[d]] for d in range(dim)]
    pi = [_snap((p_in[d] - x0[d]) / h[d]) for d in range(dim)]
    po = [_snap((p_out[d] - x0[d]) / h[d]) for d in range(dim)]
    for d in range(dim):
        assert -1e-14 <= pi[d] <= 1 + 1e-14, "current point not inside current cell"
    face, lam = -1, 100.0
    for d in range(dim):
        if po[d] < 0:
            l = (0 - pi[d]) / (po[d] - pi[d])
            if l < lam:
                lam, face = l, 2 * d
        elif po[d] > 1:
            l = (1 - pi[d]) / (po[d] - pi[d])
            if l < lam:
                lam, face = l, 2 * d + 1
    if face == -1:
        return -1, None, None
    hb = [_snap(pi[d] + lam * (po[d] - pi[d])) for d in range(dim)]
    return face, [x0[d] + hb[d] * h[d] for d in range(dim)], lam


def shape_function_values(mesh, p, nodes, cell, point):
    """SemiLagrangianTools.cpp:15-64: snapped unit point, tensor-product Lagrange values,
    local index lexicographic (x fastest)."""
    dim = mesh.dim
    xi = []
    for d in range(dim):
        x0 = mesh.verts[d][cell[d]]
        h = mesh.verts[d][cell[d] + 1] - x0
        v = _snap((point[d] - x0) / h)
        assert 0.0 <= v <= 1.0, "departure point not inside the cell it was found in"
        xi.append(v)
    l = [lagrange_values(nodes, xi[d]) for d in range(dim)]
    if dim == 2:
        return np.array([l[0][a] * l[1][b] for b in range(p + 1) for a in range(p + 1)])
    return np.array([l[0][a] * l[1][b] * l[2][c]
                     for c in range(p + 1) for b in range(p + 1) for a in range(p + 1)])


def assemble_semilagrangian(mesh, p, e, dt, numbering=None, opposite=None):
    """Restates fillSparseObject(false).  ``e``: (Q, dim) scaled directions.
    Returns dict {(bi, bj): scipy.sparse.csr_matrix(N, N)} for the non-empty blocks."""
    dim, Q = mesh.dim, e.shape[0]
    dofs = DofMap(mesh, p, numbering)
    nodes = gll_nodes(p)
    N = dofs.N
    minus_dt_e = -dt * e
    rows = {}
    tracked = np.zeros(N, dtype=bool)
    hits = []            # m_boundaryHandler.addHit(el, bi) in call order (SemiLagrangian.cpp:381-383)
    cell_rank = {c: k for k, c in enumerate(mesh.cells())}

    def add(bi, bj, r, c, v):
        rows.setdefault((bi, bj), []).append((r, c, v))

    for cell in mesh.cells():
        local = dofs.cell_dofs(cell)
        found = {}
        not_found = deque()
        # create Lagrangian support points (SemiLagrangian.cpp:214-242)
        for il, gi in enumerate(local):
            if tracked[gi]:
                continue
            tracked[gi] = True
            if dim == 2:
                loc = (il % (p + 1), il // (p + 1))
            else:
                loc = (il % (p + 1), (il // (p + 1)) % (p + 1), il // ((p + 1) ** 2))
            x_i = [mesh.verts[d][cell[d]] + nodes[loc[d]] * (mesh.verts[d][cell[d] + 1] - mesh.verts[d][cell[d]])
                   for d in range(dim)]
            for alpha in range(1, Q):
                x_dep = [x_i[d] + minus_dt_e[alpha, d] for d in range(dim)]
                not_found.append(dict(dest=gi, dest_dir=alpha, cur_dir=alpha, dep=x_dep, cur=list(x_i),
                                      cell=tuple(cell), life=0))
        # follow paths (SemiLagrangian.cpp:246-445)
        while not_found:
            el = not_found[0]
            el["life"] += 1
            if el["life"] > 50:
                add(el["dest_dir"] - 1, el["dest_dir"] - 1, el["dest"], el["dest"], 1.0)
                not_found.popleft()
                continue
            face, pb = _face_crossed_first(mesh, el["cell"], el["cur"], el["dep"])
            if face == -1:
                found.setdefault(el["cell"], []).append(el)
                not_found.popleft()
                continue
            d, side = face // 2, face % 2
            c = el["cell"]
            at_boundary = (c[d] == 0 and side == 0) or (c[d] == mesh.n[d] - 1 and side == 1)
            if at_boundary:
                if mesh.boundary[d] == "periodic":
                    L = mesh.verts[d][-1] - mesh.verts[d][0]
                    shift = L if side == 0 else -L
                    cur = list(pb)
                    cur[d] += shift
                    dep = list(el["dep"])
                    dep[d] += shift
                    nc = list(c)
                    nc[d] = mesh.n[d] - 1 if side == 0 else 0
                    el["cur"], el["dep"], el["cell"] = cur, dep, tuple(nc)
                else:
                    # VelocityNeqBounceBack: reverse direction, reflect remaining path (lines 358-384)
                    assert opposite is not None
                    el["cur_dir"] = int(opposite[el["cur_dir"]])
                    el["cur"] = list(pb)
                    ed = e[el["cur_dir"]]
                    vel = math.sqrt(float(np.dot(ed, ed)))
                    dist = math.sqrt(sum((el["dep"][k] - el["cur"][k]) ** 2 for k in range(dim)))
                    el["dep"] = [el["cur"][k] - ed[k] * dist / vel for k in range(dim)]
                    # BoundaryHit ctor (BoundaryHit.h:40-56): dtHit = dt - |currentPoint - departurePoint| / |e_cur|
                    hits.append(dict(index=el["dest"], direction=el["dest_dir"], point=tuple(el["cur"]), cell=cell_rank[tuple(c)],
                                     boundary=(d, side), dt_hit=dt - dist / vel, seq=len(hits)))
            else:
                nc = list(c)
                nc[d] += -1 if side == 0 else 1
                el["cur"], el["cell"] = list(pb), tuple(nc)
        # assemble (SemiLagrangian.cpp:450-504)
        for src_cell, lst in found.items():
            src_dofs = dofs.cell_dofs(src_cell)
            for el in lst:
                vals = shape_function_values(mesh, p, nodes, src_cell, el["dep"])
                for j, v in enumerate(vals):
                    if abs(v) < 1e-10:
                        continue
                    add(el["dest_dir"] - 1, el["cur_dir"] - 1, el["dest"], src_dofs[j], float(v))

    blocks = {}
    for key, trip in rows.items():
        r = np.array([t[0] for t in trip], dtype=np.int64)
        c = np.array([t[1] for t in trip], dtype=np.int64)
        v = np.array([t[2] for t in trip], dtype=np.float64)
        m = sp.coo_matrix((v, (r, c)), shape=(N, N)).tocsr()   # add() accumulates duplicates
        m.sort_indices()
        blocks[key] = m
    # HitList iteration order of SemiLagrangianBoundaryHandler::operate (SemiLagrangianBoundaryHandler.cpp:43-109):
    # std::map over cells, then over hit points, then the hits of a point in insertion order
    dofs.hits = sorted(hits, key=lambda h: (h["cell"], h["point"], h["seq"]))
    return blocks, dofs
