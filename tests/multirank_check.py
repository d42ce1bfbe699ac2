"""Multi-GPU parity worker (run under torchrun, one rank per GPU): slab-partitioned problem, NCCL ghost
exchange inside nb200_step, result gathered on rank 0 and compared with the single-domain CPU oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from natrium_b200 import Context, harness, _capi  # noqa: E402
from natrium_b200.stencils import Stencil          # noqa: E402


def run_case(rank, world, local, uid, name, dim, cells, p, scaling, nu, cfl, with_g, steps=6, grid=False, walls=None, host_chunks=0, blocks=None):
    """grid: the host numbers its DoFs cell by cell and gives the structure hint (TMA box kernels, ghost values also land in
    the grid copies); otherwise lexicographic numbering + the internal order hint (staged kernels).
    walls: per-axis flags -> ThermalBounceBack(0.85) walls + EXACT_DIFFERENCE force (D3Q45 f+g): the reference's order
    stream(f) -> hits on f and g -> exchange g -> stream(g) -> collide across ranks.
    host_chunks > 0: the steps go through nb200_step_host with pinned host buffers and that many pipeline pieces (uploads of the
    pieces that hold send values first, exchange behind them, rows that read ghosts behind the exchange)."""
    st = Stencil(name, scaling)
    verts = None
    if walls is not None:
        y = np.arange(cells[1] + 1) / cells[1]
        verts = [np.linspace(0, 2.0, cells[0] + 1), 1.0 * (y - 0.8 * np.sin(2 * np.pi * y) / (2 * np.pi)), np.linspace(0, 2.0, cells[2] + 1)]
    pb = harness.CartesianProblem(dim, cells, p, verts=verts)
    dt = pb.timestep(st, cfl)
    # blocks: a block partition (2 x 1 x 1 on 2 ranks cuts ACROSS the x-fastest numbering, 2 x 2 x 1 / 2 x 2 x 2 on 4 / 8 ranks
    # give 3 / 6 neighbours with edge ghosts) instead of slabs along the last axis
    part = harness.BlockPartition(pb, st, dt, rank, blocks) if blocks is not None else harness.SlabPartition(pb, st, dt, rank, world)
    ctx = Context(local, rank, world, uid)
    ctx.set_stencil(st.getDirections(), st.getWeights(), st.getScaling(), st.getSpeedOfSoundSquare())
    ctx.set_layout(part.n_owned, part.n_ghost, with_g)
    num = harness.CellNumbering(part) if grid else None
    host = num if grid else part
    if grid:
        ctx.set_dof_grid(*host.grid_coords(), fe_order=p)
    else:
        ctx.set_dof_order(part.cell_blocked_order())      # internal order: halo indices are translated inside
    F = np.array([2e-2, 0.0, 0.0])
    hits = None
    if walls is not None:
        _, hi, hd = harness.upload_streaming_matrix_walled(ctx, pb, part, st, dt, walls, num)
        hits = (hi, hd, np.full(len(hi), _capi.WALL_THERMAL_BOUNCE_BACK, dtype=np.int32), np.full(len(hi), 0.85))
        ctx.set_wall_hits(*hits)
    else:
        harness.upload_streaming_matrix(ctx, pb, part, st, dt, num)
    ctx.set_halo(*host.halo_plan())
    if walls is not None:
        ctx.set_collision(nu, dt, equilibrium=_capi.QUARTIC_EQUILIBRIUM, with_g=True, gamma=1.4, prandtl=0.7, sutherland=True,
                          force=F, force_type=_capi.EXACT_DIFFERENCE)
    elif with_g:
        ctx.set_collision(nu, dt, equilibrium=_capi.QUARTIC_EQUILIBRIUM, with_g=True, gamma=1.4, prandtl=0.71, sutherland=True)
    else:
        ctx.set_collision(nu, dt)
    # a rank that asserted here would leave its peers waiting in the exchange: the kernels in use are gathered and judged on rank 0
    grid_in_use = int(ctx.grid_info()["in_use"]) if grid else -1
    x = host.owned_points()
    if dim == 2:
        rho, u = harness.taylor_green_2d(x)
        rho = 1.0 + 0.05 * np.cos(x[:, 0]) * np.sin(x[:, 1])
        u = 0.1 * u
    elif with_g:      # low-Mach compressible start (the quartic equilibrium is not meant for |u| ~ 1 at cs2 = 1/3)
        rho, u = harness.taylor_green_3d(x, st.getSpeedOfSound(), compressible=True, density_numerator=0.1)
        u = 0.1 * u
    else:
        rho, u = harness.taylor_green_3d(x, st.getSpeedOfSound())
    T = 1.0 + 0.02 * np.sin(x[:, 0]) * np.cos(x[:, -1])
    if walls is not None:
        T = 0.85 + 0.1 * np.sin(np.pi * x[:, 1])
    if with_g:
        f, g = harness.quartic_equilibrium_distributions(st, rho, u, T, 1.4)
        ctx.upload_populations(1, g)
    else:
        f = harness.equilibrium_distributions(st, rho, u)
    ctx.upload_populations(0, f)
    if host_chunks:
        n_loc, Qn, Dn = host.n_owned if grid else part.n_owned, st.getQ(), st.getD()
        bufs = [torch.empty((Qn, n_loc), dtype=torch.float64, pin_memory=True) for _ in range(2)]
        mom = torch.empty((1 + Dn, n_loc), dtype=torch.float64, pin_memory=True)
        bufs[0].numpy()[...] = f
        l0 = ctx.kernel_launches()
        for s_ in range(steps):
            ctx.step_host(bufs[s_ & 1].data_ptr(), bufs[(s_ + 1) & 1].data_ptr(), mom.data_ptr(), mom.data_ptr() + 8 * n_loc, host_chunks)
        ctx.synchronize()
        assert ctx.kernel_launches() - l0 >= steps * host_chunks, "the pipelined path did not run"
        assert np.array_equal(ctx.download_populations(0), bufs[steps & 1].numpy())
    else:
        ctx.step(steps)
    ctx.synchronize()
    cons = ctx.conserved()
    got = [ctx.download_populations(0)] + ([ctx.download_populations(1)] if with_g else [])
    ids = part.owned_global_ids()
    if grid:
        ids = ids[num.order]
    gathered = [None] * world
    dist.all_gather_object(gathered, (ids, got, f, g if with_g else None, grid_in_use))
    ctx.close()
    if rank != 0:
        return True
    # single-domain oracle on the assembled global state
    from oracle import cpu
    import scipy.sparse as sp
    N = pb.N
    F0, G0 = np.empty((st.getQ(), N)), np.empty((st.getQ(), N))
    RES = [np.empty((st.getQ(), N)) for _ in got]
    for ids_r, got_r, f_r, g_r, _ in gathered:
        F0[:, ids_r] = f_r
        if with_g:
            G0[:, ids_r] = g_r
        for k, a in enumerate(got_r):
            RES[k][:, ids_r] = a
    single = harness.SlabPartition(pb, st, dt, 0, 1)
    blocks = {}
    ost = cpu.Stencil(name, scaling)
    if walls is not None:
        hi_all, hd_all = [], []
        for a in range(1, st.getQ()):
            bl, rh = harness.assemble_direction_walled(pb, single, st, dt, a, walls)
            for k, (rp, col, val) in bl.items():
                if len(val):
                    blocks[k] = sp.csr_matrix((val, col, rp), shape=(N, N))
            hi_all.append(rh.astype(np.int32)); hd_all.append(np.full(len(rh), a, dtype=np.int32))
        hi_all, hd_all = np.concatenate(hi_all), np.concatenate(hd_all)
        kinds, vals = np.ones(len(hi_all), dtype=np.int32), np.full(len(hi_all), 0.85)
        for _ in range(steps):
            F0 = cpu.stream(blocks, F0)
            assert cpu.apply_wall_hits(ost, F0, G0, hi_all, hd_all, kinds, vals) == 0
            G0 = cpu.stream(blocks, G0)
            assert cpu.collide_bgk_fg_forced(ost, F0, G0, nu, dt, F, "EXACT_DIFFERENCE", gamma=1.4, prandtl=0.7, sutherland=True)[-1] == 0
    else:
        for a in range(1, st.getQ()):
            rp, col, val = harness.assemble_direction(pb, single, st, dt, a)
            blocks[(a - 1, a - 1)] = sp.csr_matrix((val, col, rp), shape=(N, N))
        stepper = cpu.ReferenceOrderStepper(ost, blocks, N, nu, dt, equilibrium=1 if with_g else 0, with_g=with_g,
                                            gamma=1.4, prandtl=0.71 if with_g else None, sutherland=with_g)
        for _ in range(steps):
            assert stepper.step(F0, G0 if with_g else None) == 0
    err = float(np.max(np.abs(RES[0] - F0) / np.abs(F0)))
    if with_g:
        err = max(err, float(np.max(np.abs(RES[1] - G0) / np.abs(G0))))
    mass = F0.sum()
    kernels_ok = all(t[4] != 0 for t in gathered)        # grid cases: every rank really ran the TMA box kernels
    ok = err <= 1e-11 and abs(cons[0] - mass) <= 1e-12 * mass and kernels_ok
    print(f"multirank {name} world={world} {'grid' if grid else 'staged'}{' walled' if walls is not None else ''}: max rel err after {steps} steps = {err:.3e}, mass {cons[0]:.15g} vs {mass:.15g} -> {'OK' if ok else 'FAIL'}{'' if kernels_ok else ' (grid tables not in use on some rank)'}", flush=True)
    return ok


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    box = [Context.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    cases = [  # (label, args, kwargs)
        ("staged-d3q19", ("D3Q19", 3, [3, 3, 2 * world], 2, np.sqrt(3) / 0.05, 2 * np.pi, 0.4, False), {}),
        ("staged-d2q25", ("D2Q25H", 2, [5, 3 * world], 2, 1.0, 0.01, 1.0, True), {}),
        # D3Q45 f+g: the unfused path (stream f,g in one matrix pass + collide), interior / boundary CTA split included
        ("staged-d3q45", ("D3Q45", 3, [2, 2, 2 * world], 2, 1.0, 0.01, 0.4, True), dict(steps=3)),
        # the same through the grid (TMA box) kernels: ghost values land in the grid copies as well
        ("grid-d3q19", ("D3Q19", 3, [3, 3, 2 * world], 2, np.sqrt(3) / 0.05, 2 * np.pi, 0.4, False), dict(grid=True)),
        ("grid-d3q19-p4", ("D3Q19", 3, [4, 2, 2 * world], 4, np.sqrt(3) / 0.05, 2 * np.pi, 0.4, False), dict(grid=True)),
        ("grid-d2q25", ("D2Q25H", 2, [5, 3 * world], 2, 1.0, 0.01, 1.0, True), dict(grid=True)),
        ("grid-d3q45", ("D3Q45", 3, [2, 2, 2 * world], 2, 1.0, 0.01, 0.4, True), dict(steps=3, grid=True)),
        # block partition: staged kernels, exchange with every neighbour the partition has
        ("block-d3q19", ("D3Q19", 3, [4, 4, 4], 2, np.sqrt(3) / 0.05, 2 * np.pi, 0.4, False),
         dict(blocks={2: [2, 1, 1], 4: [2, 2, 1], 8: [2, 2, 2]}.get(world))),
        ("block-d2q25", ("D2Q25H", 2, [6, 6], 2, 1.0, 0.01, 1.0, True), dict(blocks={2: [2, 1], 4: [2, 2], 8: [4, 2]}.get(world))),
        # host-buffer steps across ranks: the chunk pipeline with the exchange inside (>= 2048 rows per rank to engage it)
        ("hoststep-d3q19-p4", ("D3Q19", 3, [4, 4, 2 * world], 4, np.sqrt(3) / 0.05, 2 * np.pi, 0.4, False), dict(host_chunks=4, steps=4)),
        ("hoststep-cellnum", ("D3Q19", 3, [4, 4, 2 * world], 4, np.sqrt(3) / 0.05, 2 * np.pi, 0.4, False), dict(host_chunks=3, steps=3, grid=True)),
        # walls across ranks (thermal hits rewrite g before it is exchanged and streamed), staged and grid kernels
        ("staged-walled", ("D3Q45", 3, [2, 3, 2 * world], 2, 1.0, 0.01, 0.4, True), dict(steps=3, walls=[False, True, False])),
        ("grid-walled", ("D3Q45", 3, [2, 3, 2 * world], 2, 1.0, 0.01, 0.4, True), dict(steps=3, grid=True, walls=[False, True, False])),
    ]
    only = [k for k in os.environ.get("MULTIRANK_CASES", "").split(",") if k]
    ok = True
    for label, a, kw in cases:
        if only and label not in only:
            continue
        if label.startswith("block-") and kw.get("blocks") is None:
            continue
        if rank == 0:
            print(f"case {label} ...", flush=True)
        box = [Context.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ok = run_case(rank, world, local, box[0], *a, **kw) and ok
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    dist.destroy_process_group()
    sys.exit(int(flag.item() != 0))


if __name__ == "__main__":
    main()
