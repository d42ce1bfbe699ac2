"""Host-side logic on CPU: the synthetic assembly (product harness) against the oracle's path-tracing
assembly, the stencil tables against the oracle's and the reference sources, and the slab partition /
ghost plan (single process simulation of the exchange)."""
import math
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp

from natrium_b200 import harness
from natrium_b200.stencils import Stencil
from oracle import assembly, stencils as ost

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

REF = "/root/reference/src/library/natrium/stencils"


@pytest.mark.parametrize("name", ["D2Q9", "D3Q19", "D3Q15", "D2Q25H", "D3Q45"])
def test_stencil_tables(name):
    s = Stencil(name, 1.7)
    e, w, cs2, vm = ost.make(name, 1.7)
    assert np.array_equal(s.getDirections(), e) and np.array_equal(s.getWeights(), w)
    assert s.getSpeedOfSoundSquare() == cs2 and s.getMaxParticleVelocityMagnitude() == vm
    assert abs(w.sum() - 1) < 1e-12
    assert np.max(np.abs(w @ e)) < 1e-12                          # first moment vanishes
    second = np.einsum("i,ia,ib->ab", w, e, e)
    assert np.max(np.abs(second - cs2 * np.eye(e.shape[1]))) < 1e-10   # isotropy with cs2 = scaling^2/3
    for i in range(s.getQ()):
        j = s.getIndexOfOppositeDirection(i)
        assert np.max(np.abs(e[i] + e[j])) < 1e-14


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference sources not mounted (GPU box)")
def test_d3q45_table_against_reference_source():
    """The 45 literal direction rows and weights, parsed straight from L/stencils/D3Q45.cpp."""
    src = open(os.path.join(REF, "D3Q45.cpp")).read()
    rows = re.findall(r"\{\s*(-?[0-9.]+)\s*,\s*(-?[0-9.]+)\s*,\s*(-?[0-9.]+)\s*\}", src)
    raw = np.array([[float(a) for a in r] for r in rows])
    assert raw.shape == (45, 3)
    e, w, _, _ = ost.make("D3Q45", 2.0)
    assert np.array_equal(e, 2.0 * raw / math.sqrt(3))
    wsrc = re.search(r"vector<double> result \{(.*?)\};", src, flags=re.S).group(1)
    wref = np.array([float(x) for x in re.findall(r"[0-9.]+", wsrc)])
    assert np.array_equal(w, wref)


def test_gll_points():
    for p in range(1, 9):
        a, b = harness.gauss_lobatto_points(p), assembly.gll_nodes(p)
        assert np.max(np.abs(a - b)) < 2e-16
    assert abs(harness.gauss_lobatto_points(4)[1] - (1 - math.sqrt(3 / 7)) / 2) < 1e-16


CASES = [(2, 8, 4, "D2Q9", math.sqrt(3) / 0.05, 0.4, 1), (2, 4, 2, "D2Q25H", 1.0, 1.0, 1), (2, 6, 3, "D2Q9", 1.0, 0.4, 3),
         (3, 3, 2, "D3Q19", math.sqrt(3) / 0.05, 0.4, 1), (3, 4, 2, "D3Q19", 1.0, 0.4, 2), (3, 2, 2, "D3Q45", 1.0, 0.4, 1),
         (3, 4, 1, "D3Q15", 1.0, 0.4, 4)]


def local_to_global(pb, part):
    gid = part.owned_global_ids()
    if len(part.ghost_planes):
        gid = np.concatenate([gid, (part.ghost_planes[:, None] * pb.plane + np.arange(pb.plane)[None, :]).reshape(-1)])
    return gid


@pytest.mark.parametrize("dim,cells,p,name,scaling,cfl,nranks", CASES)
def test_harness_assembly_equals_oracle(dim, cells, p, name, scaling, cfl, nranks):
    """The vectorised product-side assembly reproduces the oracle's path-tracing restatement of
    fillSparseObject entry by entry, for every rank of a slab partition."""
    st = Stencil(name, scaling)
    pb = harness.CartesianProblem(dim, cells, p)
    dt = pb.timestep(st, cfl)
    e, w, cs2, vm = ost.make(name, scaling)
    mesh = assembly.CartesianMesh.uniform(dim, cells)
    assert dt == assembly.calculate_timestep(mesh, p, vm, cfl)
    blocks, dofs = assembly.assemble_semilagrangian(mesh, p, e, dt)
    assert all(k[0] == k[1] for k in blocks)                 # periodic: diagonal blocks only
    seen = np.zeros(pb.N, dtype=int)
    for r in range(nranks):
        part = harness.SlabPartition(pb, st, dt, r, nranks)
        own, gid = part.owned_global_ids(), local_to_global(pb, part)
        seen[own] += 1
        for a in range(1, st.getQ()):
            rp, col, val = harness.assemble_direction(pb, part, st, dt, a)
            coo = sp.csr_matrix((val, col, rp), shape=(part.n_owned, part.n_owned + part.n_ghost)).tocoo()
            mg = sp.coo_matrix((coo.data, (coo.row, gid[coo.col])), shape=(part.n_owned, pb.N)).tocsr()
            ref = blocks[(a - 1, a - 1)][own]
            assert mg.nnz == ref.nnz
            d = mg - ref
            assert d.nnz == 0 or np.max(np.abs(d.data)) <= 1e-15
    assert np.all(seen == 1)                                  # every DoF owned exactly once


@pytest.mark.parametrize("nranks", [2, 3, 4])
def test_halo_plan_single_process(nranks):
    """Simulate the exchange of nb200_set_halo's plan in numpy: after it, rank-local SpMVs equal the global one."""
    st = Stencil("D3Q19", 1.0)
    pb = harness.CartesianProblem(3, [2, 2, 4], 2)
    dt = pb.timestep(st, 0.4)
    parts = [harness.SlabPartition(pb, st, dt, r, nranks) for r in range(nranks)]
    plans = [p.halo_plan() for p in parts]
    rng = np.random.default_rng(0)
    xg = rng.standard_normal(pb.N)
    local = [np.concatenate([xg[p.owned_global_ids()], np.full(p.n_ghost, np.nan)]) for p in parts]
    for r, (nbr, so, si, ro) in enumerate(plans):
        assert ro[-1] == parts[r].n_ghost
        for k, dst in enumerate(nbr):
            nb2, so2, si2, ro2 = plans[dst]
            k2 = list(nb2).index(r)                           # the matching segment on the receiver
            payload = local[r][si[so[k]:so[k + 1]]]
            assert len(payload) == ro2[k2 + 1] - ro2[k2]
            local[dst][parts[dst].n_owned + ro2[k2]: parts[dst].n_owned + ro2[k2 + 1]] = payload
    single = harness.SlabPartition(pb, st, dt, 0, 1)
    for a in (1, 2, 7, 16):
        rp, col, val = harness.assemble_direction(pb, single, st, dt, a)
        yg = sp.csr_matrix((val, col, rp), shape=(pb.N, pb.N)) @ xg
        for r, p in enumerate(parts):
            assert not np.isnan(local[r]).any()
            rp, col, val = harness.assemble_direction(pb, p, st, dt, a)
            y = sp.csr_matrix((val, col, rp), shape=(p.n_owned, p.n_owned + p.n_ghost)) @ local[r]
            assert np.max(np.abs(y - yg[p.owned_global_ids()])) <= 1e-14


def test_equilibrium_init_against_oracle():
    from oracle import fields
    st = Stencil("D3Q19", 3.0)
    e, w, cs2, _ = ost.make("D3Q19", 3.0)
    rng = np.random.default_rng(1)
    rho, u = 1 + 0.1 * rng.standard_normal(20), 0.2 * rng.standard_normal((3, 20))
    assert np.max(np.abs(harness.equilibrium_distributions(st, rho, u) - fields.equilibrium_init(e, w, cs2, rho, u))) <= 1e-15
    for name in ("D2Q25H", "D3Q45"):
        st = Stencil(name, 1.3)
        e, w, cs2, _ = ost.make(name, 1.3)
        D = e.shape[1]
        rho, u, T = 1 + 0.1 * rng.standard_normal(7), 0.2 * rng.standard_normal((D, 7)), 1 + 0.1 * rng.standard_normal(7)
        f1, g1 = harness.quartic_equilibrium_distributions(st, rho, u, T, 1.4)
        f2, g2 = fields.quartic_equilibrium_init(e, w, cs2, 1.3, rho, u, T, 1.4)
        assert np.max(np.abs(f1 - f2) / np.abs(f2)) <= 1e-11 and np.max(np.abs(g1 - g2) / np.abs(g2)) <= 1e-11


def test_staging_tables_replay(tmp_path):
    """Host-side builder of the staged driving tables (natrium_b200/csrc/dict_build.h): a CPU replay of the kernel's
    access pattern over the tables equals the CSR product, passes respect their capacity and cover every
    direction once, and an unshareable matrix is reported infeasible (the library then keeps the plain kernels)."""
    import subprocess
    exe = str(tmp_path / "staging_check")
    src = os.path.join(ROOT, "tests", "cpp", "staging_check.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-o", exe, src], check=True)
    for args, expect in [(("1000", "4", "16", "9", "4096", "1"), "OK"), (("5000", "18", "64", "25", "600", "3"), "OK"),
                         (("300000", "6", "32", "5", "256", "4"), "OK"), (("777", "3", "1", "40", "4096", "5"), "INFEASIBLE")]:
        out = subprocess.run([exe, *args], check=True, capture_output=True, text=True).stdout
        assert out.startswith(expect), out


def test_grid_tables_replay(tmp_path):
    """Host-side builder of the grid (TMA box) driving tables (natrium_b200/csrc/grid_build.h): a CPU replay of the kernel's
    access pattern -- boxes of the lexicographic grid copy (out-of-range points read 0, everything outside the boxes is
    poisoned), rows from (offset + per-direction offset table), generic rows from their dictionary list -- equals the CSR
    product for 2-d / 3-d periodic tensor grids, lexicographic and random DoF numbering, shuffled row entries, FE orders
    1-4, and walls (off-diagonal entries and truncated rows must come out generic)."""
    import subprocess
    exe = str(tmp_path / "grid_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-o", exe, os.path.join(ROOT, "tests", "cpp", "grid_check.cpp")], check=True)
    for args in ["3 3 3 3 4 0 1536 1", "3 3 2 4 4 1 1536 2", "2 5 4 0 4 1 1536 3", "3 4 3 3 2 1 1536 4", "3 2 3 2 3 1 1536 5",
                 "2 4 4 0 2 0 1536 6 1", "3 3 3 2 4 1 1536 7 1", "3 2 2 2 1 0 1536 8", "3 6 5 4 4 1 1280 9", "2 9 7 0 4 1 1536 11"]:
        out = subprocess.run([exe, *args.split()], check=True, capture_output=True, text=True).stdout
        assert out.startswith("OK"), (args, out)
    # the harness' own grid coordinates: one grid point per local DoF, cell faces at multiples of p, for every rank
    from natrium_b200 import harness
    from natrium_b200.stencils import Stencil
    st = Stencil("D3Q19", 3.0)
    pb = harness.CartesianProblem(3, [2, 2, 8], 4)
    dt = pb.timestep(st, 0.4)
    for r in range(4):
        part = harness.SlabPartition(pb, st, dt, r, 4)
        dims, coords = part.grid_coords()
        assert coords.shape == (part.n_owned + part.n_ghost, 3) and (coords >= 0).all() and (coords < dims[None, :]).all()
        assert len(np.unique(coords, axis=0)) == len(coords)
        num = harness.CellNumbering(part)
        d2, c2 = num.grid_coords()
        assert np.array_equal(c2[:part.n_owned], coords[:part.n_owned][num.order]) and np.array_equal(c2[part.n_owned:], coords[part.n_owned:])


def _dump_rank_problem(path, name, scaling, dim, cells, p, cfl, rank, world, cell_numbering, partition=None):
    """One rank of a partitioned harness problem in the file format of tests/cpp/grid_check.cpp (file mode)."""
    from natrium_b200 import harness
    st = Stencil(name, scaling)
    pb = harness.CartesianProblem(dim, cells, p)
    dt = pb.timestep(st, cfl)
    part = (partition or harness.SlabPartition)(pb, st, dt, rank, world)
    num = harness.CellNumbering(part) if cell_numbering else None
    dims, coords = (num if cell_numbering else part).grid_coords()
    with open(path, "wb") as f:
        f.write(np.array([dim, p, part.n_owned, part.n_ghost, st.getQ() - 1] + list(dims) + [1] * (3 - dim), dtype=np.int64).tobytes())
        f.write(np.ascontiguousarray(coords, dtype=np.int32).tobytes())
        for a in range(1, st.getQ()):
            rp, col, val = harness.assemble_direction(pb, part, st, dt, a)
            if num is not None:
                rp, col, val = num.renumber_csr(rp, col, val)
            f.write(np.array([len(val)], dtype=np.int64).tobytes())
            f.write(np.ascontiguousarray(rp, dtype=np.int64).tobytes())
            f.write(np.ascontiguousarray(col, dtype=np.int32).tobytes())
            f.write(np.ascontiguousarray(val, dtype=np.float64).tobytes())
    return part


def test_grid_tables_replay_partitioned(tmp_path):
    """The grid builder on the ranks of partitioned problems (ghost slots in the grid, rows whose lists straddle the periodic
    seam or the ghost layer): the tables must be feasible on every rank (the multi-GPU bench runs the TMA box kernels), replay
    to the CSR product, and leave only a thin layer of rows to the dictionary lists.  Regression: rank 0 of a 2-slab p = 4
    mesh used to pick a seam-straddling list as the direction's template and gave up."""
    import subprocess
    exe = str(tmp_path / "grid_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-o", exe, os.path.join(ROOT, "tests", "cpp", "grid_check.cpp")], check=True)
    cases = [("D3Q19", math.sqrt(3) / 0.05, 3, [4, 2, 4], 4, 0.4, 2, 1536), ("D3Q19", math.sqrt(3) / 0.05, 3, [3, 3, 6], 2, 0.4, 3, 1536),
             ("D2Q25H", 1.0, 2, [5, 6], 2, 1.0, 2, 1536), ("D3Q45", 1.0, 3, [2, 2, 4], 2, 0.4, 2, 1280)]
    for name, sc, dim, cells, p, cfl, world, cap in cases:
        for r in range(world):
            for cn in (True, False):
                path = str(tmp_path / "case.bin")
                _dump_rank_problem(path, name, sc, dim, cells, p, cfl, r, world, cn)
                out = subprocess.run([exe, "file", path, str(cap)], check=True, capture_output=True, text=True).stdout
                assert out.startswith("OK"), (name, cells, r, world, cn, out)
                kv = dict(t.split("=") for t in out.split()[1:])
                assert int(kv["box_rows"]) >= 0.85 * int(kv["checked"]), out


@pytest.mark.parametrize("name,scaling,boundary,p,cfl,stretch", [
    ("D2Q9", 3.0, ["periodic", "wall"], 3, 0.8, True), ("D2Q25H", 1.0, ["wall", "wall"], 2, 1.0, False),
    ("D3Q45", 1.0, ["periodic", "wall", "periodic"], 2, 0.4, True), ("D3Q19", 2.0, ["periodic", "wall", "periodic"], 3, 0.4, True),
    ("D2Q9", 2.0, ["wall", "wall"], 4, 0.4, True)])
def test_walled_assembly_equals_oracle(name, scaling, boundary, p, cfl, stretch):
    """harness.assemble_direction_walled (vectorised input generator for the walled bench configurations: Riemann 2D with
    walls all around, the channel with walls in y) against the oracle's restatement of fillSparseObject with bounce-back
    walls (path reversal at the wall, off-diagonal blocks, double bounce and stuck paths in corners, one hit per bounce)."""
    import scipy.sparse as sp
    from oracle import assembly, cpu
    from natrium_b200 import harness
    from natrium_b200.stencils import Stencil
    ost, st = cpu.Stencil(name, scaling), Stencil(name, scaling)
    opp = harness.opposite_directions(st)
    dim = len(boundary)
    y = np.linspace(0, 1, 4)
    ys = y - 0.8 * np.sin(2 * np.pi * y) / (2 * np.pi) if stretch else np.linspace(0, 2, 5)
    verts = [np.linspace(0, 1, 3) if stretch else np.linspace(0, 2, 5), ys] + ([np.linspace(0, 1, 4)] if dim == 3 else [])
    mesh = assembly.CartesianMesh(verts, boundary=boundary)
    dt = assembly.calculate_timestep(mesh, p, ost.max_speed, cfl)
    blocks, dofs = assembly.assemble_semilagrangian(mesh, p, ost.e, dt, opposite=opp)
    pb = harness.CartesianProblem(dim, [len(v) - 1 for v in verts], p, verts=verts)
    assert pb.timestep(st, cfl) == dt
    part = harness.SlabPartition(pb, st, dt)
    mine, hits = {}, set()
    for a in range(1, st.getQ()):
        bl, rh = harness.assemble_direction_walled(pb, part, st, dt, a, [b == "wall" for b in boundary], opp)
        for k, (rp, c, v) in bl.items():
            m = sp.csr_matrix((v, c, rp), shape=(part.n_owned, part.n_owned))
            if m.nnz:
                mine[k] = m
        hits |= {(int(r), a) for r in rh}
    assert set(mine) == set(blocks)
    assert max(abs(mine[k] - blocks[k]).max() for k in blocks) <= 1e-14
    assert hits == {(h["index"], h["direction"]) for h in dofs.hits}


def test_cell_numbering_renumbers_consistently():
    """harness.CellNumbering (host numbering = deal.II-like cell-wise order): P A P^T applied to P x equals P (A x),
    entry order inside a row (= summation order) is untouched, ghost columns keep their slots, the halo plan follows."""
    st = Stencil("D3Q19", math.sqrt(3) / 0.05)
    pb = harness.CartesianProblem(3, [3, 2, 4], 2)
    dt = pb.timestep(st, 0.4)
    for rank, world in [(0, 1), (1, 2)]:
        part = harness.SlabPartition(pb, st, dt, rank, world)
        num = harness.CellNumbering(part)
        n, ng = part.n_owned, part.n_ghost
        assert sorted(num.order.tolist()) == list(range(n))
        rng = np.random.default_rng(0)
        x = rng.standard_normal(n + ng)
        xp = np.concatenate([x[:n][num.order], x[n:]])
        for alpha in (1, 7, 12):
            rp, col, val = harness.assemble_direction(pb, part, st, dt, alpha)
            rp2, col2, val2 = num.renumber_csr(rp, col, val)
            A = sp.csr_matrix((val, col, rp), shape=(n, n + ng))
            B = sp.csr_matrix((val2, col2, rp2), shape=(n, n + ng))
            assert np.max(np.abs(B @ xp - (A @ x)[num.order])) <= 1e-14
            assert np.array_equal(val2.reshape(n, -1), val.reshape(n, -1)[num.order])
        assert np.allclose(num.owned_points(), part.owned_points()[num.order])
        if world > 1:
            nbr, so, si, ro = part.halo_plan()
            nbr2, so2, si2, ro2 = num.halo_plan()
            assert np.array_equal(num.order[si2], si) and np.array_equal(so, so2) and np.array_equal(ro, ro2)


def test_dictionary_builder_semantics(tmp_path):
    """Host-side dictionary builder (natrium_b200/csrc/dict_build.h): tolerance 0 dedups bitwise only and replays the
    CSR product exactly; the default 1e-14 merges round-off-noisy copies of a weight pattern (error <= K * tol) but
    keeps values 1e-9 apart separate; more distinct row lengths than exact classes fall into power-of-two classes
    with zero-weight padding; rows fed by two blocks of a block-row (wall bounce) are concatenated."""
    import subprocess
    exe = str(tmp_path / "dict_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-o", exe, os.path.join(ROOT, "tests", "cpp", "dict_check.cpp")], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    assert out.startswith("OK"), out


def test_filter_level_schedule(tmp_path):
    """Host-side level schedule of the exponential filter (natrium_b200/csrc/filter_build.h): cells of a level share no DoF,
    cells that share a DoF keep their order across levels, and a non-commuting cell update applied level by level (any order
    inside a level) equals the sequential loop bit for bit -- for lexicographic, reversed, Morton and shuffled cell orders."""
    import subprocess
    exe = str(tmp_path / "filter_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "cpp", "filter_check.cpp")], check=True)
    for args in ["2 4 3 1 2 0 1", "2 6 5 1 3 1 2", "3 4 4 4 2 2 3", "3 3 4 2 4 3 4", "3 8 8 8 1 0 5", "2 9 7 1 4 2 6", "3 16 16 16 1 2 7"]:
        out = subprocess.run([exe, *args.split()], check=True, capture_output=True, text=True).stdout
        assert out.startswith("OK"), (args, out)


def test_grid_tables_replay_walled(tmp_path):
    """Walled meshes through the grid builder (multi-block file mode of tests/cpp/grid_check.cpp: bounce blocks included in the
    replayed product): Riemann-like D2Q25H p = 2 with walls on all four sides and a y-walled D3Q19 p = 2 channel.  Regression:
    the corner DoF of a walled mesh has a one-entry row (stuck path) and used to make that row length "class 0" for the 8
    directions that point into a corner -- a third of all rows then left the TMA boxes (BASELINE config 4).  Now class 0 is the
    class most rows have and only the wall layer is taken from the dictionary lists."""
    import subprocess
    from natrium_b200 import harness
    exe = str(tmp_path / "grid_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-o", exe, os.path.join(ROOT, "tests", "cpp", "grid_check.cpp")], check=True)
    cases = [("D2Q25H", 1.0, 2, [32, 32], 2, 1.0, [True, True], [2.0, 2.0], 512, 0.95),
             ("D3Q19", 2.0, 3, [4, 6, 3], 2, 0.4, [False, True, False], [2.0, 1.0, 2.0], 1536, 0.80)]
    for name, sc, dim, cells, p, cfl, walls, length, cap, min_share in cases:
        st = Stencil(name, sc)
        pb = harness.CartesianProblem(dim, cells, p, length=length)
        dt = pb.timestep(st, cfl)
        part = harness.SlabPartition(pb, st, dt)
        num = harness.CellNumbering(part)
        dims, coords = num.grid_coords()
        path = str(tmp_path / "walled.bin")
        with open(path, "wb") as f:
            f.write(np.array([dim, p, part.n_owned, part.n_ghost, -(st.getQ() - 1)] + list(dims) + [1] * (3 - dim), dtype=np.int64).tobytes())
            f.write(np.ascontiguousarray(coords, dtype=np.int32).tobytes())
            for a in range(1, st.getQ()):
                bl, _ = harness.assemble_direction_walled(pb, part, st, dt, a, walls)
                items = [(k, v) for k, v in bl.items() if len(v[2])]
                f.write(np.array([len(items)], dtype=np.int64).tobytes())
                for (bi, bj), (rp, col, val) in items:
                    rp, col, val = num.renumber_csr(rp, col, val)
                    f.write(np.array([bj, len(val)], dtype=np.int64).tobytes())
                    f.write(np.ascontiguousarray(rp, dtype=np.int64).tobytes())
                    f.write(np.ascontiguousarray(col, dtype=np.int32).tobytes())
                    f.write(np.ascontiguousarray(val, dtype=np.float64).tobytes())
        out = subprocess.run([exe, "file", path, str(cap)], check=True, capture_output=True, text=True).stdout
        assert out.startswith("OK"), (name, out)
        kv = dict(t.split("=") for t in out.split()[1:])
        assert int(kv["box_rows"]) >= min_share * int(kv["checked"]), (name, out)


def test_grid_row_pairs_are_unified_within_the_tolerance(tmp_path):
    """The kernels multiply rows t and t + 64 of a tile together when their pattern ids agree; a mismatch drags the whole warp
    through both code paths.  At the bench's cell size (32 cells over 2 pi) round-off puts 2-5 % of those pairs -- same
    position in x-neighbour cells, weights equal up to 1.3e-14 -- into neighbouring tolerance buckets.  The grid builder gives
    such a pair one pattern (twice the value tolerance): >= 99.9 % paired afterwards, products still equal to the CSR product
    (replay), nothing unified with tolerance 0."""
    import subprocess
    exe = str(tmp_path / "grid_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-o", exe, os.path.join(ROOT, "tests", "cpp", "grid_check.cpp")], check=True)
    from natrium_b200 import harness
    st = Stencil("D3Q19", math.sqrt(3) / 0.05)
    cells, length = [32, 4, 2], [2 * math.pi, 2 * math.pi / 8, 2 * math.pi / 16]
    pb = harness.CartesianProblem(3, cells, 4, length=length)
    dt = pb.timestep(st, 0.4)
    part = harness.SlabPartition(pb, st, dt)
    num = harness.CellNumbering(part)
    dims, coords = num.grid_coords()
    path = str(tmp_path / "u32.bin")
    with open(path, "wb") as f:
        f.write(np.array([3, 4, part.n_owned, part.n_ghost, 18] + list(dims), dtype=np.int64).tobytes())
        f.write(np.ascontiguousarray(coords, dtype=np.int32).tobytes())
        for a in range(1, 19):
            rp, col, val = num.renumber_csr(*harness.assemble_direction(pb, part, st, dt, a))
            f.write(np.array([len(val)], dtype=np.int64).tobytes())
            f.write(np.ascontiguousarray(rp, dtype=np.int64).tobytes())
            f.write(np.ascontiguousarray(col, dtype=np.int32).tobytes())
            f.write(np.ascontiguousarray(val, dtype=np.float64).tobytes())
    res = {}
    for tol in ("0", "1e-14"):
        out = subprocess.run([exe, "file", path, "1536"], check=True, capture_output=True, text=True, env=dict(os.environ, GRID_CHECK_TOL=tol)).stdout
        assert out.startswith("OK"), out
        res[tol] = dict(t.split("=") for t in out.split()[1:])
    assert int(res["0"]["unified"]) == 0
    assert int(res["1e-14"]["unified"]) > 0 and float(res["1e-14"]["paired"]) >= 0.999
