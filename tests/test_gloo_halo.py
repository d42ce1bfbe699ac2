"""World-size-2 (and 3) CPU test of the N>1 host logic over torch.distributed/gloo: every rank builds its
slab partition (or, 4 / 8 ranks, its block of a 2 x 2 x 1 / 2 x 2 x 2 block partition: up to 6 neighbours with edge ghosts,
the shape of a p4est Z-curve partition) and ghost plan, exchanges ghosts with isend/irecv exactly as nb200_set_halo's plan prescribes
(one message per neighbour, [population][entry] payload), multiplies with its local CSR blocks and the
gathered result must equal the single-domain product."""
import os

import numpy as np
import pytest
import scipy.sparse as sp
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from natrium_b200 import harness
from natrium_b200.stencils import Stencil


def _worker(rank, world, port, out, blocks=None):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    st = Stencil("D3Q19", 1.0)
    pb = harness.CartesianProblem(3, [2, 3, 2 * world] if blocks is None else [4, 4, 2 * blocks[2]], 2)
    dt = pb.timestep(st, 0.4)
    part = harness.SlabPartition(pb, st, dt, rank, world) if blocks is None else harness.BlockPartition(pb, st, dt, rank, blocks)
    nbr, so, si, ro = part.halo_plan()
    Q = st.getQ()
    rng = np.random.default_rng(5)
    xg = rng.standard_normal((Q, pb.N))                      # same on every rank
    x = np.concatenate([xg[:, part.owned_global_ids()], np.full((Q, part.n_ghost), np.nan)], axis=1)
    reqs, recv_bufs = [], []
    for k, r in enumerate(nbr):
        send = torch.from_numpy(np.ascontiguousarray(x[1:, si[so[k]:so[k + 1]]]))      # streamed populations only
        buf = torch.empty((Q - 1, int(ro[k + 1] - ro[k])), dtype=torch.float64)
        recv_bufs.append(buf)
        reqs.append(dist.isend(send, int(r)))
        reqs.append(dist.irecv(buf, int(r)))
    for q in reqs:
        q.wait()
    for k in range(len(nbr)):
        x[1:, part.n_owned + ro[k]: part.n_owned + ro[k + 1]] = recv_bufs[k].numpy()
    y = np.zeros((Q, part.n_owned))
    for a in range(1, Q):
        rp, col, val = harness.assemble_direction(pb, part, st, dt, a)
        y[a] = sp.csr_matrix((val, col, rp), shape=(part.n_owned, part.n_owned + part.n_ghost)) @ x[a]
    gathered = [None] * world
    dist.all_gather_object(gathered, (part.owned_global_ids(), y, len(nbr)))
    if rank == 0:
        if blocks is not None:
            assert max(g[2] for g in gathered) >= (3 if world == 4 else 6)      # more than a slab's two neighbours
        single = harness.SlabPartition(pb, st, dt, 0, 1)
        worst = 0.0
        Y = np.zeros((Q, pb.N))
        for ids, yr, _ in gathered:
            Y[:, ids] = yr
        for a in range(1, Q):
            rp, col, val = harness.assemble_direction(pb, single, st, dt, a)
            ref = sp.csr_matrix((val, col, rp), shape=(pb.N, pb.N)) @ xg[a]
            worst = max(worst, float(np.max(np.abs(Y[a] - ref))))
        out.put(worst)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_ghost_exchange_matches_single_domain(world):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29600 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    worst = out.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert worst <= 1e-13


@pytest.mark.parametrize("blocks", [[2, 2, 1], [2, 2, 2]])
def test_gloo_block_partition_exchange_matches_single_domain(blocks):
    world = int(np.prod(blocks))
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29610 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, out, blocks)) for r in range(world)]
    for p in procs:
        p.start()
    worst = out.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert worst <= 1e-13
