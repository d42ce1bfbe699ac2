"""Synthetic host side: what deal.II/p4est + SemiLagrangian::fillSparseObject hand to the GPU.

In a real NATriuM build the mesh, DoFHandler and the assembled streaming matrix come from
the reference's host C++.  This module produces the same *kind* of input for
tests and benchmarks on machines without deal.II: Cartesian (optionally stretched) periodic
meshes, continuous FE_Q(p) on Gauss-Lobatto points, the per-direction semi-Lagrangian
interpolation matrix as local CSR blocks, a slab partition with its ghost plan, and the
Taylor-Green initial fields.  Everything is vectorised numpy so the 129^3-DoF / 7e8-nnz
benchmark matrix is built in about a minute, block by block.

It is an input generator, not part of the accelerated path, and it is independent of
``oracle/`` (tests check one against the other).

What it produces follows (L = src/library/natrium):
  matrix entries   L/advection/SemiLagrangian.cpp:201-242,476-501 (N_j(x_i - dt e_alpha), drop < 1e-10)
  unit-cell snap   L/advection/SemiLagrangianTools.cpp:41-47
  dt               L/utilities/CFDSolverUtilities.cpp:92-100
  TGV fields       L/benchmarks/TaylorGreenVortex2D.cpp:36-60, TaylorGreenVortex3D.cpp:38-72
  f = f_eq init    L/solver/CFDSolver.cpp:1104-1129, L/solver/CompressibleCFDSolver.h:790-901
"""
import math

import numpy as np


# ------------------------------------------------------------------------------------------
# 1D building blocks
# ------------------------------------------------------------------------------------------
def gauss_lobatto_points(p):
    """Gauss-Lobatto-Legendre nodes on [0,1] (p+1 of them)."""
    if p == 1:
        return np.array([0.0, 1.0])
    # eigenvalues of the Jacobi matrix for the interior nodes (roots of P'_p = Jacobi(1,1) of degree p-1)
    k = np.arange(1, p - 1, dtype=np.float64)
    beta = np.sqrt(k * (k + 2) / ((2 * k + 1) * (2 * k + 3)))
    J = np.diag(beta, 1) + np.diag(beta, -1)
    x = np.linalg.eigvalsh(J) if p > 2 else np.array([0.0])
    # polish with Newton on q(x) = (1-x^2) P_p'(x)
    for _ in range(5):
        P0, P1 = np.ones_like(x), x.copy()
        for n in range(2, p + 1):
            P0, P1 = P1, ((2 * n - 1) * x * P1 - (n - 1) * P0) / n
        x = x - (p * (P0 - x * P1)) / (-p * (p + 1) * P1)
    x = np.sort(np.concatenate([[-1.0], x, [1.0]]))
    x = 0.5 * (x - x[::-1])          # exact antisymmetry
    return 0.5 * (x + 1.0)


def lagrange_basis(nodes, xi):
    """All p+1 Lagrange polynomials at points xi: shape (len(xi), p+1); product form."""
    xi = np.asarray(xi, dtype=np.float64)
    n = len(nodes)
    out = np.ones((xi.shape[0], n))
    for j in range(n):
        for m in range(n):
            if m != j:
                out[:, j] *= (xi - nodes[m]) / (nodes[j] - nodes[m])
    return out


def _snap01(v):
    v = np.where(np.abs(v) < 1e-10, 0.0, v)
    return np.where(np.abs(v - 1.0) < 1e-10, 1.0, v)


class Axis:
    """One coordinate axis: cell vertices, the continuous 1D DoF grid and departure tracking."""

    def __init__(self, verts, p, nodes):
        self.v = np.asarray(verts, dtype=np.float64)
        self.n = len(self.v) - 1
        self.p = p
        self.nodes = nodes
        self.nd = self.n * p + 1
        g = np.arange(self.nd)
        self.home = np.where(g > 0, (g - 1) // p, 0)          # first cell (lexicographic) that sees the DoF
        self.loc = g - self.home * p
        h = self.v[self.home + 1] - self.v[self.home]
        self.x = self.v[self.home] + nodes[self.loc] * h

    def track(self, delta):
        """Departure of every 1D DoF displaced by ``delta`` through periodic cells.
        Returns (cols, wts): per DoF the k 1D DoF indices / weights with k = 1 if delta == 0 else p+1."""
        p, n, v = self.p, self.n, self.v
        L = v[-1] - v[0]
        c = self.home.copy()
        xd = self.x + delta
        for _ in range(60):
            xi = _snap01((xd - v[c]) / (v[c + 1] - v[c]))
            lo, hi = xi < 0, xi > 1
            if not (lo.any() or hi.any()):
                break
            wrap_lo, wrap_hi = lo & (c == 0), hi & (c == n - 1)
            xd = np.where(wrap_lo, xd + L, np.where(wrap_hi, xd - L, xd))
            c = np.where(lo, np.where(c == 0, n - 1, c - 1), np.where(hi, np.where(c == n - 1, 0, c + 1), c))
        else:
            raise RuntimeError("departure point tracking did not terminate")
        W = lagrange_basis(self.nodes, xi)
        if delta == 0.0:
            j = np.argmax(np.abs(W), axis=1)
            cols = (c * p + j)[:, None]
            wts = np.take_along_axis(W, j[:, None], axis=1)
        else:
            cols = c[:, None] * p + np.arange(p + 1)[None, :]
            wts = W
        return cols.astype(np.int64), wts


# ------------------------------------------------------------------------------------------
# problem + partition
# ------------------------------------------------------------------------------------------
class CartesianProblem:
    """Periodic hyper-rectangle, cells[d] cells per axis, FE order p.  DoFs are numbered
    lexicographically (x fastest); periodic faces are not identified (as in the reference)."""

    def __init__(self, dim, cells, p, length=2 * math.pi, verts=None):
        self.dim, self.p = dim, p
        cells = [cells] * dim if np.isscalar(cells) else list(cells)
        length = [length] * dim if np.isscalar(length) else list(length)
        if verts is None:
            verts = [length[d] * np.arange(cells[d] + 1) / cells[d] for d in range(dim)]
        self.nodes = gauss_lobatto_points(p)
        self.axes = [Axis(verts[d], p, self.nodes) for d in range(dim)]
        self.nd = [a.nd for a in self.axes]
        self.N = int(np.prod(self.nd))
        self.plane = int(np.prod(self.nd[:-1]))       # DoFs per plane of the last axis

    def min_vertex_distance(self):
        return min(float(np.min(np.diff(a.v))) for a in self.axes)

    def timestep(self, stencil, cfl):
        return cfl * self.min_vertex_distance() / (stencil.getMaxParticleVelocityMagnitude() * self.p * self.p)

    def points_of_planes(self, planes):
        """Support points (n, dim) of all DoFs on the given last-axis planes, lexicographic."""
        ax = [a.x for a in self.axes[:-1]] + [self.axes[-1].x[np.asarray(planes)]]
        grids = np.meshgrid(*ax[::-1], indexing="ij")[::-1]   # last axis slowest, x fastest
        return np.stack([g.reshape(-1) for g in grids], axis=1)


class SlabPartition:
    """Rank r owns a contiguous range of planes of the last axis (cells split evenly; the shared
    plane between two slabs belongs to the lower rank, rank 0 also owns plane 0).  Ghosts are whole
    planes, ordered by owner rank then lexicographic id."""

    def __init__(self, problem, stencil, dt, rank=0, nranks=1):
        self.pb, self.rank, self.nranks = problem, rank, nranks
        ax = problem.axes[-1]
        if nranks > ax.n:
            raise ValueError("more ranks than cell layers")
        bounds = [(ax.n * r) // nranks for r in range(nranks + 1)]
        self.plane_owner = np.empty(ax.nd, dtype=np.int64)
        for r in range(nranks):
            lo = 0 if r == 0 else bounds[r] * ax.p + 1
            hi = bounds[r + 1] * ax.p + 1
            self.plane_owner[lo:hi] = r
        deltas = sorted(set(float(-dt * e[-1]) for e in stencil.getDirections()[1:]))
        self._tracks = {d: ax.track(d) for d in deltas}
        self.owned_planes = np.nonzero(self.plane_owner == rank)[0]
        self.ghost_planes = self._ghost_planes_of(rank)
        self.n_owned = len(self.owned_planes) * problem.plane
        self.n_ghost = len(self.ghost_planes) * problem.plane
        # plane -> local base index
        self.base = np.full(ax.nd, -1, dtype=np.int64)
        self.base[self.owned_planes] = np.arange(len(self.owned_planes)) * problem.plane
        self.base[self.ghost_planes] = self.n_owned + np.arange(len(self.ghost_planes)) * problem.plane

    def _ghost_planes_of(self, r):
        own = np.nonzero(self.plane_owner == r)[0]
        need = set()
        for cols, _ in self._tracks.values():
            need.update(np.unique(cols[own]).tolist())
        gp = np.array(sorted(need - set(own.tolist())), dtype=np.int64)
        if len(gp):
            gp = gp[np.lexsort((gp, self.plane_owner[gp]))]
        return gp

    def halo_plan(self):
        """(nbr_rank, send_off, send_idx, recv_off) for nb200_set_halo."""
        plane = self.pb.plane
        nbrs = sorted(set(self.plane_owner[self.ghost_planes].tolist())
                      | {r for r in range(self.nranks) if r != self.rank
                         and np.any(self.plane_owner[self._ghost_planes_of(r)] == self.rank)})
        send_off, recv_off, send_idx = [0], [0], []
        for r in nbrs:
            theirs = self._ghost_planes_of(r)
            mine_for_them = theirs[self.plane_owner[theirs] == self.rank]
            for pl in mine_for_them:
                send_idx.append(self.base[pl] + np.arange(plane))
            send_off.append(send_off[-1] + len(mine_for_them) * plane)
            recv_off.append(recv_off[-1] + int(np.sum(self.plane_owner[self.ghost_planes] == r)) * plane)
        send_idx = np.concatenate(send_idx).astype(np.int32) if send_idx else np.zeros(0, dtype=np.int32)
        return (np.array(nbrs, dtype=np.int32), np.array(send_off, dtype=np.int64), send_idx,
                np.array(recv_off, dtype=np.int64))

    def owned_points(self):
        return self.pb.points_of_planes(self.owned_planes)

    def cell_blocked_order(self):
        """Owned local DoF indices in the order SemiLagrangian::fillSparseObject visits them
        (SemiLagrangian.cpp:201-223): cell by cell (lexicographic, x fastest), cell->get_dof_indices(),
        first visit wins.  Input for nb200_set_dof_order."""
        pb = self.pb
        p, dim = pb.p, pb.dim
        loc = np.arange(p + 1)
        last = pb.axes[-1]
        own_cells = [c for c in range(last.n) if np.any(self.plane_owner[c * p + loc] == self.rank)]
        gz = (np.asarray(own_cells, dtype=np.int64)[:, None] * p + loc[None, :])          # (ncz, p+1) planes
        zbase = np.where(self.plane_owner[gz] == self.rank, self.base[gz], -1)
        gx = (np.arange(pb.axes[0].n, dtype=np.int64)[:, None] * p + loc[None, :])        # (ncx, p+1)
        if dim == 2:
            ids = zbase[:, None, :, None] + gx[None, :, None, :]                         # (cz, cx, lz, lx)
            ok = np.broadcast_to(zbase[:, None, :, None] >= 0, ids.shape)
        else:
            ndx = pb.nd[0]
            gy = (np.arange(pb.axes[1].n, dtype=np.int64)[:, None] * p + loc[None, :])
            ids = (zbase[:, None, None, :, None, None] + gy[None, :, None, None, :, None] * ndx
                   + gx[None, None, :, None, None, :])                                    # (cz, cy, cx, lz, ly, lx)
            ok = np.broadcast_to(zbase[:, None, None, :, None, None] >= 0, ids.shape)
        flat = ids.reshape(-1)[ok.reshape(-1)]
        uniq, first = np.unique(flat, return_index=True)
        order = uniq[np.argsort(first, kind="stable")]
        assert len(order) == self.n_owned
        return order.astype(np.int32)

    def owned_global_ids(self):
        return (self.owned_planes[:, None] * self.pb.plane + np.arange(self.pb.plane)[None, :]).reshape(-1)

    def cell_dofs(self):
        """[n_cells, (p+1)^dim] local DoF indices (owned or ghost slots) of the rank's cells, cells in lexicographic order
        (x fastest), element-local numbering lexicographic: what cell->get_dof_indices gives in the active-cell loop of a
        Cartesian deal.II mesh with an FE_DGQ-like local numbering.  Input for nb200_set_filter."""
        pb, p, dim = self.pb, self.pb.p, self.pb.dim
        last = pb.axes[-1]
        bounds = [(last.n * r) // self.nranks for r in range(self.nranks + 1)]
        loc = np.arange(p + 1, dtype=np.int64)
        cz = np.arange(bounds[self.rank], bounds[self.rank + 1], dtype=np.int64)
        zb = self.base[cz[:, None] * p + loc[None, :]]                                # (ncz, p+1) plane bases
        assert (zb >= 0).all(), "a plane of an own cell is neither owned nor a ghost"
        gx = np.arange(pb.axes[0].n, dtype=np.int64)[:, None] * p + loc[None, :]      # (ncx, p+1)
        if dim == 2:
            ids = zb[:, None, :, None] + gx[None, :, None, :]                         # (cz, cx, lz, lx)
        else:
            gy = (np.arange(pb.axes[1].n, dtype=np.int64)[:, None] * p + loc[None, :]) * pb.nd[0]
            ids = zb[:, None, None, :, None, None] + gy[None, :, None, None, :, None] + gx[None, None, :, None, None, :]
        return np.ascontiguousarray(ids.reshape(-1, (p + 1) ** dim), dtype=np.int32)

    def grid_coords(self):
        """(dims, coords) for nb200_set_dof_grid: integer grid coordinates of every local DoF (owned, then ghosts) in a
        local tensor grid whose coordinates 0, p, 2p, ... are cell faces.  In a deal.II build the same numbers come from
        ranking the support-point coordinates per axis.  The last axis is the slab axis: own cells keep their order, the
        cell above the slab follows directly, cells reached across the periodic boundary sit one empty cell away (the two
        coincident planes there are distinct DoFs)."""
        pb, p = self.pb, self.pb.p
        ax = pb.axes[-1]
        C = ax.n
        bounds = [(C * r) // self.nranks for r in range(self.nranks + 1)]
        b0, b1 = bounds[self.rank], bounds[self.rank + 1]
        vcell = {}                                     # plane -> (virtual cell, position)
        for g in self.owned_planes:
            c = 0 if g == 0 else (g - 1) // p
            vcell[int(g)] = (c, int(g) - c * p)
        for g in self.ghost_planes:
            g = int(g)
            cands = [min(g // p, C - 1)]
            if g % p == 0 and g > 0 and g // p <= C - 1:
                cands.append(g // p - 1)
            pick = None
            for c in cands:                            # a cell of this rank, or the neighbour right above / below
                if b0 <= c < b1:
                    pick = (c, g - c * p)
            if pick is None:
                for c in cands:
                    if c == b1 or c == b0 - 1:
                        pick = (c, g - c * p)
            if pick is None:                           # across the periodic boundary: the top cell below / the bottom cell above
                below = cands[0] >= b1
                c = max(cands) if below else min(cands)
                pick = ((b0 - 2) if below else (b1 + 1), g - c * p)
            vcell[g] = pick
        vmin = min(v for v, _ in vcell.values())
        zloc = {g: (v - vmin) * p + j for g, (v, j) in vcell.items()}
        nz = max(zloc.values()) + 1
        planes = np.concatenate([self.owned_planes, self.ghost_planes]).astype(np.int64)
        zl = np.array([zloc[int(g)] for g in planes], dtype=np.int32)
        if pb.dim == 2:
            x = np.arange(pb.nd[0], dtype=np.int32)
            coords = np.stack([np.tile(x, len(planes)), np.repeat(zl, pb.plane)], axis=1)
            dims = [pb.nd[0], nz]
        else:
            x = np.arange(pb.nd[0], dtype=np.int32)
            y = np.arange(pb.nd[1], dtype=np.int32)
            xy = np.stack([np.tile(x, pb.nd[1]), np.repeat(y, pb.nd[0])], axis=1)
            coords = np.concatenate([np.tile(xy, (len(planes), 1)), np.repeat(zl, pb.plane)[:, None]], axis=1)
            dims = [pb.nd[0], pb.nd[1], nz]
        assert len(np.unique(coords, axis=0)) == len(coords)
        return np.array(dims, dtype=np.int32), np.ascontiguousarray(coords, dtype=np.int32)


class BlockPartition:
    """px x py (x pz) blocks of cells -- what a p4est Z-curve partition of a regular mesh gives on 4 or 8 ranks
    (SemiLagrangian.cpp:185-193 iterates the locally owned cells of such a partition): up to 3^dim - 1 neighbours, edge and
    corner ghosts.  A DoF belongs to the rank that owns the first cell (lexicographic) that sees it, axis by axis, so the
    owned set is a box of the DoF grid; local order of the owned DoFs is lexicographic inside the box (x fastest), ghosts
    follow grouped by owner rank (ascending), inside a rank by global id."""

    def __init__(self, problem, stencil, dt, rank, blocks):
        self.pb, self.rank, self.blocks = problem, rank, list(blocks)
        dim = problem.dim
        assert len(self.blocks) == dim
        self.nranks = int(np.prod(self.blocks))
        self._stencil, self._dt = stencil, dt
        # per-axis owner of every 1-d DoF index
        self.axis_owner = []
        for d in range(dim):
            ax = problem.axes[d]
            if self.blocks[d] > ax.n:
                raise ValueError("more blocks than cells along an axis")
            bounds = [(ax.n * r) // self.blocks[d] for r in range(self.blocks[d] + 1)]
            cell_owner = np.zeros(ax.n, dtype=np.int64)
            for r in range(self.blocks[d]):
                cell_owner[bounds[r]:bounds[r + 1]] = r
            self.axis_owner.append(cell_owner[ax.home])
        self._tracks = [{float(-dt * e[d]): problem.axes[d].track(float(-dt * e[d])) for e in stencil.getDirections()[1:]} for d in range(dim)]
        self.own_axes = self._own_axes(rank)
        self.owned_gids = self._box_ids(self.own_axes)
        self.n_owned = len(self.owned_gids)
        self.ghost_gids = self._ghosts_of(rank)
        self.n_ghost = len(self.ghost_gids)
        self.g2l = np.full(problem.N, -1, dtype=np.int64)
        self.g2l[self.owned_gids] = np.arange(self.n_owned)
        self.g2l[self.ghost_gids] = self.n_owned + np.arange(self.n_ghost)

    # rank <-> block coordinates, x fastest
    def _coords_of(self, r):
        out = []
        for d in range(self.pb.dim):
            out.append(r % self.blocks[d])
            r //= self.blocks[d]
        return out

    def _own_axes(self, r):
        rc = self._coords_of(r)
        return [np.nonzero(self.axis_owner[d] == rc[d])[0] for d in range(self.pb.dim)]

    def _box_ids(self, idx):
        """global ids of the tensor product of per-axis index arrays, lexicographic (x fastest)"""
        nd = self.pb.nd
        if self.pb.dim == 2:
            return (idx[1][:, None] * nd[0] + idx[0][None, :]).reshape(-1)
        return ((idx[2][:, None, None] * nd[1] + idx[1][None, :, None]) * nd[0] + idx[0][None, None, :]).reshape(-1)

    def owner_of(self, gids):
        nd, dim = self.pb.nd, self.pb.dim
        g = np.asarray(gids, dtype=np.int64)
        ix = [g % nd[0], (g // nd[0]) % nd[1]] + ([g // (nd[0] * nd[1])] if dim == 3 else [])
        r, mul = np.zeros_like(g), 1
        for d in range(dim):
            r = r + self.axis_owner[d][ix[d]] * mul
            mul *= self.blocks[d]
        return r

    def _ghosts_of(self, r):
        own = self._own_axes(r)
        read = []
        for e in self._stencil.getDirections()[1:]:
            per_axis = [np.unique(self._tracks[d][float(-self._dt * e[d])][0][own[d]]) for d in range(self.pb.dim)]
            read.append(self._box_ids(per_axis))
        read = np.unique(np.concatenate(read))
        gh = read[self.owner_of(read) != r]
        return gh[np.lexsort((gh, self.owner_of(gh)))]

    def halo_plan(self):
        """(nbr_rank, send_off, send_idx, recv_off) for nb200_set_halo."""
        mine = self.owner_of(self.ghost_gids)
        theirs = {r: self._ghosts_of(r) for r in range(self.nranks) if r != self.rank}
        needs_me = {r: g[self.owner_of(g) == self.rank] for r, g in theirs.items()}
        nbrs = sorted(set(mine.tolist()) | {r for r, g in needs_me.items() if len(g)})
        send_off, recv_off, send_idx = [0], [0], []
        for r in nbrs:
            send_idx.append(self.g2l[needs_me[r]])
            send_off.append(send_off[-1] + len(needs_me[r]))
            recv_off.append(recv_off[-1] + int(np.sum(mine == r)))
        send_idx = np.concatenate(send_idx).astype(np.int32) if send_idx else np.zeros(0, dtype=np.int32)
        return (np.array(nbrs, dtype=np.int32), np.array(send_off, dtype=np.int64), send_idx, np.array(recv_off, dtype=np.int64))

    def owned_global_ids(self):
        return self.owned_gids

    def owned_points(self):
        ax = self.pb.axes
        grids = np.meshgrid(*[ax[d].x[self.own_axes[d]] for d in range(self.pb.dim)][::-1], indexing="ij")[::-1]
        return np.stack([g.reshape(-1) for g in grids], axis=1)

    def cell_blocked_order(self):
        """Owned local DoF indices cell by cell (own cells lexicographic, first visit wins): nb200_set_dof_order input."""
        pb, p, dim = self.pb, self.pb.p, self.pb.dim
        rc = self._coords_of(self.rank)
        loc = np.arange(p + 1, dtype=np.int64)
        cells = []
        for d in range(dim):
            n = pb.axes[d].n
            cells.append(np.arange((n * rc[d]) // self.blocks[d], (n * (rc[d] + 1)) // self.blocks[d], dtype=np.int64))
        g = [cells[d][:, None] * p + loc[None, :] for d in range(dim)]         # (cells_d, p+1) 1-d DoF indices
        nd = pb.nd
        if dim == 2:
            ids = g[1][:, None, :, None] * nd[0] + g[0][None, :, None, :]
        else:
            ids = (g[2][:, None, None, :, None, None] * nd[1] + g[1][None, :, None, None, :, None]) * nd[0] + g[0][None, None, :, None, None, :]
        loc_ids = self.g2l[ids.reshape(-1)]
        loc_ids = loc_ids[(loc_ids >= 0) & (loc_ids < self.n_owned)]
        uniq, first = np.unique(loc_ids, return_index=True)
        order = uniq[np.argsort(first, kind="stable")]
        assert len(order) == self.n_owned
        return order.astype(np.int32)

    def dense_direction(self, alpha):
        """(col, val), shape (owned rows, k): columns in local numbering, entries in (z, y, x) order"""
        pb, dim = self.pb, self.pb.dim
        e = self._stencil.getDirection(alpha)
        tr = [self._tracks[d][float(-self._dt * e[d])] for d in range(dim)]
        c = [tr[d][0][self.own_axes[d]] for d in range(dim)]
        w = [tr[d][1][self.own_axes[d]] for d in range(dim)]
        nd = pb.nd
        if dim == 2:
            col = c[1][:, None, :, None] * nd[0] + c[0][None, :, None, :]
            val = w[0][None, :, None, :] * w[1][:, None, :, None]
        else:
            col = (c[2][:, None, None, :, None, None] * nd[1] + c[1][None, :, None, None, :, None]) * nd[0] + c[0][None, None, :, None, None, :]
            val = (w[0][None, None, :, None, None, :] * w[1][None, :, None, None, :, None]) * w[2][:, None, None, :, None, None]
        k = int(np.prod(col.shape[dim:]))
        rows = int(np.prod(col.shape[:dim]))
        col = self.g2l[np.broadcast_to(col, val.shape).reshape(rows, k)]
        if (col < 0).any():
            raise RuntimeError("a row reads a DoF that is neither owned nor a ghost")
        return np.ascontiguousarray(col), np.ascontiguousarray(val.reshape(rows, k))


# ------------------------------------------------------------------------------------------
# streaming matrix
# ------------------------------------------------------------------------------------------
def _dense_direction(problem, part, stencil, dt, alpha):
    """(col, val) of shape (owned rows, k): the periodic tensor-product row of every owned DoF, entries in (z, y, x) order."""
    if isinstance(part, BlockPartition):
        return part.dense_direction(alpha)
    e = stencil.getDirection(alpha)
    dim = problem.dim
    tr = [problem.axes[d].track(float(-dt * e[d])) for d in range(dim)]
    own = part.owned_planes
    zc, zw = tr[-1][0][own], tr[-1][1][own]
    zbase = part.base[zc]
    if (zbase < 0).any():
        raise RuntimeError("departure point outside owned + ghost planes (time step too large)")
    if dim == 2:
        xc, xw = tr[0]
        col = zbase[:, None, :, None] + xc[None, :, None, :]
        val = xw[None, :, None, :] * zw[:, None, :, None]
    else:
        xc, xw = tr[0]
        yc, yw = tr[1]
        ndx = problem.nd[0]
        col = (zbase[:, None, None, :, None, None] + yc[None, :, None, None, :, None] * ndx
               + xc[None, None, :, None, None, :])
        val = (xw[None, None, :, None, None, :] * yw[None, :, None, None, :, None]) * zw[:, None, None, :, None, None]
    k = int(np.prod(col.shape[dim:]))
    rows = int(np.prod(col.shape[:dim]))
    col = np.ascontiguousarray(np.broadcast_to(col, val.shape).reshape(rows, k))
    val = np.ascontiguousarray(val.reshape(rows, k))
    return col, val


def _csr_from_pieces(n_rows, pieces):
    """CSR (rowptr int64, col int32, val f64) from pieces (rows ascending and unique over all pieces, col (m, K), val (m, K));
    entries below the reference's 1e-10 drop threshold are left out (SemiLagrangian.cpp:483)."""
    lens = np.zeros(n_rows, dtype=np.int64)
    oks = []
    for rr, c, v in pieces:
        ok = np.abs(v) >= 1e-10
        lens[rr] = ok.sum(axis=1)
        oks.append(ok)
    rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    oc = np.empty(rp[-1], dtype=np.int32)
    ov = np.empty(rp[-1])
    for (rr, c, v), ok in zip(pieces, oks):
        if len(rr) == 0:
            continue
        k = ok.sum(axis=1)
        pos = np.repeat(rp[rr], k) + (np.arange(int(k.sum())) - np.repeat(np.cumsum(k) - k, k))
        oc[pos] = c[ok]
        ov[pos] = v[ok]
    return rp, oc, ov


def assemble_direction(problem, part, stencil, dt, alpha):
    """Local CSR (rowptr int64, col int32, val f64) of block (alpha-1, alpha-1): rows = owned DoFs
    of ``part`` in local order, columns in local numbering (owned, then ghosts)."""
    col, val = _dense_direction(problem, part, stencil, dt, alpha)
    rows, k = val.shape
    keep = np.abs(val) >= 1e-10
    if keep.all():
        rowptr = np.arange(rows + 1, dtype=np.int64) * k
        return rowptr, col.reshape(-1).astype(np.int32), val.reshape(-1)
    rowptr = np.concatenate([[0], np.cumsum(keep.sum(axis=1))]).astype(np.int64)
    return rowptr, col[keep].astype(np.int32), val[keep]


def opposite_directions(stencil):
    e = stencil.getDirections()
    return np.array([int(np.argmin(np.abs(e + e[i]).sum(1))) for i in range(len(e))])


def assemble_direction_walled(problem, part, stencil, dt, alpha, walls, opposite=None):
    """Direction ``alpha`` on a mesh whose axes with ``walls[d]`` end in bounce-back walls (VelocityNeqBounceBack /
    ThermalBounceBack; the other axes are periodic): what fillSparseObject does when a path reaches such a wall
    (L/advection/SemiLagrangian.cpp:358-384) -- the direction is reversed and the rest of the path runs back from the hit
    point.  Along the line x + t (-dt e_alpha) the walls bound t to [t_lo, t_hi] (t_lo <= 0 <= t_hi); a path of length 1
    folds at t_hi (first bounce: the row reads population opposite(alpha) at t = 2 t_hi - 1 and lands in the off-diagonal
    block) and, in a corner, once more at t_lo (second bounce: population alpha again, t = 2 t_lo - 2 t_hi + 1); one
    BoundaryHit per bounce.  Vectorised: the bounced rows are grouped by their distinct end points t (a handful: only
    the DoFs of the wall layer bounce), each group is a tensor product of 1-d tracks like the regular rows.
    Returns ({(bi, bj): (rowptr, col, val)}, hit_rows): local CSR blocks over the owned rows of ``part`` and the local
    index of the destination row of every hit (a twice-bounced row appears twice); the hit direction is alpha."""
    opposite = opposite_directions(stencil) if opposite is None else opposite
    e = stencil.getDirection(alpha)
    dim = problem.dim
    delta = [float(-dt * e[d]) for d in range(dim)]
    own = part.owned_planes
    idx = [np.arange(problem.nd[d]) for d in range(dim - 1)] + [own]
    shape = tuple(len(i) for i in idx[::-1])                         # (z, y, x) / (y, x): last axis slowest
    t_hi = np.full(shape, np.inf)
    t_lo = np.full(shape, -np.inf)
    for d in range(dim):
        if not walls[d] or delta[d] == 0.0:
            continue
        ax = problem.axes[d]
        xs = ax.x[idx[d]]
        lo, hi = ax.v[0], ax.v[-1]
        fwd, bwd = (lo, hi) if delta[d] < 0 else (hi, lo)           # wall met moving along +delta / -delta
        th = (fwd - xs) / delta[d]
        tl = (bwd - xs) / delta[d]
        th = np.where(np.abs(th) < 1e-14, 0.0, th)
        tl = np.where(np.abs(tl) < 1e-14, 0.0, tl)
        sh = [1] * dim
        sh[dim - 1 - d] = len(xs)
        t_hi = np.minimum(t_hi, th.reshape(sh))
        t_lo = np.maximum(t_lo, tl.reshape(sh))
    t_hi, t_lo = np.broadcast_to(t_hi, shape).reshape(-1), np.broadcast_to(t_lo, shape).reshape(-1)
    n_rows = int(np.prod(shape))
    # crossing: the end point t = 1 lies beyond the wall by more than the reference's snap tolerance (1e-10 of a cell)
    tol = 1e-10
    one = t_hi < 1.0 - tol
    t1 = np.where(one, 2.0 * t_hi - 1.0, 1.0)
    two = one & (t1 < t_lo - tol)
    t_end = np.where(two, 2.0 * t_lo - 2.0 * t_hi + 1.0, t1)
    # a DoF in a corner whose path leaves through one wall and whose reflection leaves through the other never gets
    # anywhere: the reference's tracker gives up after 50 rounds and puts 1.0 on the diagonal (SemiLagrangian.cpp:252-268)
    stuck = one & (t_hi - t_lo < tol)
    assert not (two & ~stuck & (t_end > t_hi + tol)).any(), "third wall hit"
    one_sp = one & ~stuck
    dense_c, dense_v = _dense_direction(problem, part, stencil, dt, alpha)
    diag = [(np.nonzero(~one)[0], dense_c[~one], dense_v[~one])]
    if stuck.any():
        rs = np.nonzero(stuck)[0]
        diag.append((rs, rs[:, None].astype(np.int64), np.ones((len(rs), 1))))
    off = []
    special = np.nonzero(one_sp)[0]
    coords = np.unravel_index(special, shape)[::-1]               # per axis (x, y[, z]) positions into idx[d]
    t_sp = t_end[special]
    for t in np.unique(t_sp):
        sel = t_sp == t
        rr = special[sel]
        tr = [problem.axes[d].track(delta[d] * t) for d in range(dim)]
        gl = [idx[d][coords[d][sel]] for d in range(dim)]            # global 1-d DoF index per axis
        zc, zw = tr[-1][0][gl[-1]], tr[-1][1][gl[-1]]
        zbase = part.base[zc]
        if (zbase < 0).any():
            raise RuntimeError("reflected departure point outside owned + ghost planes")
        xc, xw = tr[0][0][gl[0]], tr[0][1][gl[0]]
        if dim == 2:
            c = zbase[:, :, None] + xc[:, None, :]
            v = zw[:, :, None] * xw[:, None, :]
        else:
            yc, yw = tr[1][0][gl[1]], tr[1][1][gl[1]]
            ndx = problem.nd[0]
            c = zbase[:, :, None, None] + yc[:, None, :, None] * ndx + xc[:, None, None, :]
            v = (xw[:, None, None, :] * yw[:, None, :, None]) * zw[:, :, None, None]
        m = len(rr)
        c = np.broadcast_to(c, v.shape).reshape(m, -1)
        v = v.reshape(m, -1)
        tw = two[rr]
        if tw.any():
            diag.append((rr[tw], c[tw], v[tw]))
        if (~tw).any():
            off.append((rr[~tw], c[~tw], v[~tw]))
    blocks = {(alpha - 1, alpha - 1): _csr_from_pieces(n_rows, diag)}
    if off:
        blocks[(alpha - 1, int(opposite[alpha]) - 1)] = _csr_from_pieces(n_rows, off)
    hit_rows = np.sort(np.concatenate([np.nonzero(one)[0], np.nonzero(two & ~stuck)[0]]))
    return blocks, hit_rows


def upload_streaming_matrix_walled(ctx, problem, part, stencil, dt, walls, numbering=None):
    """upload_streaming_matrix for a mesh with bounce-back walls.  Returns (nnz, hit_index, hit_direction): the hit list
    in the caller's numbering (one hit per bounced path), for nb200_set_wall_hits."""
    nnz = 0
    hi, hd = [], []
    opp = opposite_directions(stencil)
    for alpha in range(1, stencil.getQ()):
        blocks, rows_hit = assemble_direction_walled(problem, part, stencil, dt, alpha, walls, opp)
        for (bi, bj), (rowptr, col, val) in blocks.items():
            if numbering is not None:
                rowptr, col, val = numbering.renumber_csr(rowptr, col, val)
            ctx.upload_block_csr(bi, bj, rowptr, col, val)
            nnz += len(val)
        idx = rows_hit if numbering is None else numbering.perm[rows_hit]
        hi.append(np.asarray(idx, dtype=np.int32))
        hd.append(np.full(len(rows_hit), alpha, dtype=np.int32))
    ctx.finalize_matrix()
    return nnz, np.concatenate(hi), np.concatenate(hd)


class CellNumbering:
    """The same partition with the owned DoFs numbered the way deal.II numbers them: cell by cell in the order
    fillSparseObject walks (SlabPartition.cell_blocked_order).  With it the host's numbering already is the one the
    staged kernels like, no nb200_set_dof_order hint is needed, and host buffers map 1:1 onto device arrays."""

    def __init__(self, part):
        self.part = part
        self.order = part.cell_blocked_order().astype(np.int64)      # new index k <- old index order[k]
        self.perm = np.empty_like(self.order)
        self.perm[self.order] = np.arange(len(self.order))
        self.n_owned, self.n_ghost = part.n_owned, part.n_ghost

    def owned_points(self):
        return self.part.owned_points()[self.order]

    def grid_coords(self):
        dims, coords = self.part.grid_coords()
        n = self.n_owned
        return dims, np.ascontiguousarray(np.concatenate([coords[:n][self.order], coords[n:]]))

    def halo_plan(self):
        nbr, send_off, send_idx, recv_off = self.part.halo_plan()
        return nbr, send_off, self.perm[send_idx].astype(np.int32), recv_off

    def cell_dofs(self):
        """SlabPartition.cell_dofs in this numbering (ghost slots keep their indices)."""
        cd = self.part.cell_dofs()
        n = self.part.n_owned
        out = cd.copy()
        m = cd < n
        out[m] = self.perm[cd[m]]
        return out

    def renumber_csr(self, rowptr, col, val):
        n = self.n_owned
        lens = np.diff(rowptr)
        if len(val) and np.all(lens == lens[0]):
            k = int(lens[0])
            col2 = col.reshape(n, k)[self.order].reshape(-1)
            val2 = val.reshape(n, k)[self.order].reshape(-1)
            rowptr2 = rowptr
        else:
            new_lens = lens[self.order]
            rowptr2 = np.concatenate([[0], np.cumsum(new_lens)]).astype(np.int64)
            src = np.repeat(rowptr[:-1][self.order] - rowptr2[:-1], new_lens) + np.arange(rowptr2[-1])
            col2, val2 = col[src], val[src]
        owned = col2 < n
        col2 = np.where(owned, self.perm[np.where(owned, col2, 0)], col2).astype(np.int32)
        return rowptr2, col2, val2


def row_noise(val, alpha, eps):
    """Worst-case input for the device format: every stored value gets its own relative perturbation (deterministic,
    |.| <= eps), so that no two rows of the matrix share a weight pattern -- what an unstructured mesh gives.  Applied to the
    lexicographic-numbering CSR of direction alpha, before any renumbering, by the product and the checker alike."""
    k = np.arange(len(val), dtype=np.uint64) + np.uint64(alpha) * np.uint64(0x9E3779B97F4A7C15)
    k ^= k >> np.uint64(31); k *= np.uint64(0xBF58476D1CE4E5B9); k ^= k >> np.uint64(29)
    u = (k >> np.uint64(11)).astype(np.float64) / float(1 << 53)              # [0, 1)
    return val * (1.0 + eps * (2.0 * u - 1.0))


def upload_streaming_matrix(ctx, problem, part, stencil, dt, numbering=None, noise=0.0):
    """Assemble and hand over all diagonal blocks one at a time (peak host memory = one block).
    numbering: a CellNumbering of ``part`` -- rows and owned columns are renumbered before the upload.
    noise > 0: row_noise on every block (worst-case matrix for the device format)."""
    nnz = 0
    for alpha in range(1, stencil.getQ()):
        rowptr, col, val = assemble_direction(problem, part, stencil, dt, alpha)
        if noise > 0.0:
            val = row_noise(val, alpha, noise)
        if numbering is not None:
            rowptr, col, val = numbering.renumber_csr(rowptr, col, val)
        ctx.upload_block_csr(alpha - 1, alpha - 1, rowptr, col, val)
        nnz += len(val)
    ctx.finalize_matrix()
    return nnz


# ------------------------------------------------------------------------------------------
# initial fields
# ------------------------------------------------------------------------------------------
def taylor_green_2d(x, length=2 * math.pi):
    k = 2 * math.pi / length
    u = np.stack([np.sin(k * x[:, 0]) * np.cos(k * x[:, 1]), -np.cos(k * x[:, 0]) * np.sin(k * x[:, 1])])
    return np.ones(x.shape[0]), u


def taylor_green_3d(x, cs, compressible=False, density_numerator=1.0):
    u = np.stack([np.sin(x[:, 0]) * np.cos(x[:, 1]) * np.cos(x[:, 2]),
                  -np.cos(x[:, 0]) * np.sin(x[:, 1]) * np.cos(x[:, 2]), np.zeros(x.shape[0])])
    pr = (np.cos(2 * x[:, 0]) + np.cos(2 * x[:, 1])) * (np.cos(2 * x[:, 2]) + 2) / 16.
    rho = 1.0 + (pr * density_numerator if compressible else pr / (cs * cs))
    return rho, u


def equilibrium_distributions(stencil, rho, u):
    """f_eq(rho, u) to second order (BGKStandard::getEquilibriumDistribution); returns (Q, n)."""
    e, w, cs2 = stencil.getDirections(), stencil.getWeights(), stencil.getSpeedOfSoundSquare()
    eu = e @ u / cs2
    uu = np.sum(u * u, axis=0) / (2 * cs2)
    return w[:, None] * rho[None, :] * (1 + eu * (1 + 0.5 * eu) - uu[None, :])


def quartic_equilibrium_distributions(stencil, rho, u, T, gamma):
    """Fourth-order Hermite equilibrium with temperature and g = f_eq T (2 Cv - D)
    (CompressibleCFDSolver::initializeDistributions); tensor form via einsum.  Returns f, g (Q, n)."""
    s = stencil.getScaling()
    e = stencil.getDirections() / s
    w = stencil.getWeights()
    cs2 = stencil.getSpeedOfSoundSquare() / (s * s)
    D = e.shape[1]
    v = u / s
    I = np.eye(D)
    T1 = cs2 * (T - 1.0)
    # Hermite tensors of the directions
    H2 = np.einsum("ia,ib->iab", e, e) - cs2 * I
    H3 = (np.einsum("ia,ib,ic->iabc", e, e, e)
          - cs2 * (np.einsum("ia,bc->iabc", e, I) + np.einsum("ib,ac->iabc", e, I) + np.einsum("ic,ab->iabc", e, I)))
    ee = np.einsum("ia,ib->iab", e, e)
    H4 = (np.einsum("ia,ib,ic,id->iabcd", e, e, e, e)
          - cs2 * (np.einsum("iab,cd->iabcd", ee, I) + np.einsum("iac,bd->iabcd", ee, I) + np.einsum("iad,bc->iabcd", ee, I)
                   + np.einsum("ibc,ad->iabcd", ee, I) + np.einsum("ibd,ac->iabcd", ee, I) + np.einsum("icd,ab->iabcd", ee, I))
          + cs2 * cs2 * (np.einsum("ab,cd->abcd", I, I) + np.einsum("ac,bd->abcd", I, I) + np.einsum("ad,bc->abcd", I, I)))
    # moments of the Maxwellian
    a2 = np.einsum("an,bn->abn", v, v) + I[:, :, None] * T1
    a3 = (np.einsum("an,bn,cn->abcn", v, v, v)
          + T1 * (np.einsum("ab,cn->abcn", I, v) + np.einsum("bc,an->abcn", I, v) + np.einsum("ac,bn->abcn", I, v)))
    vv = np.einsum("an,bn->abn", v, v)
    dd = np.einsum("ab,cd->abcd", I, I) + np.einsum("ac,bd->abcd", I, I) + np.einsum("ad,bc->abcd", I, I)
    a4 = (np.einsum("an,bn,cn,dn->abcdn", v, v, v, v)
          + T1 * (np.einsum("abn,cd->abcdn", vv, I) + np.einsum("acn,bd->abcdn", vv, I) + np.einsum("adn,bc->abcdn", vv, I)
                  + np.einsum("bcn,ad->abcdn", vv, I) + np.einsum("bdn,ac->abcdn", vv, I) + np.einsum("cdn,ab->abcdn", vv, I))
          + T1 * T1 * dd[..., None])
    series = (1.0 + (e @ v) / cs2
              + np.einsum("iab,abn->in", H2, a2) / (2 * cs2 ** 2)
              + np.einsum("iabc,abcn->in", H3, a3) / (6 * cs2 ** 3)
              + np.einsum("iabcd,abcdn->in", H4, a4) / (24 * cs2 ** 4))
    f = w[:, None] * rho[None, :] * series
    g = f * T * (2.0 / (gamma - 1.0) - D)
    return f, g
