#!/usr/bin/env python
"""bench.py -- throughput of the fused stream+collide step (BASELINE.json metric).

  python bench.py --gpus 1 --steps K --warmup W            # our arm (CUDA, through the C ABI)
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   # weak scaling, one rank per GPU
  python bench.py --impl reference ...                     # the reference's CPU path on the host cores

The line (rank 0, one JSON object on stdout):
  value / ms_per_step   BASELINE.json configs[1] -- Taylor-Green vortex 3D, D3Q19 BGK, semi-Lagrangian streaming, FE order 4,
                        refinement 5 (32^3 cells, 129^3 = 2 146 689 DoFs, 7.08e8 matrix non-zeros) per GPU, everything
                        resident in HBM; N GPUs hold N such slabs stacked along z (weak scaling, NCCL ghost exchange).
  e2e                   the same step driven with HOST buffers through the C ABI (pinned H2D of f, step, D2H of f, rho, u).
  roofline              the dominant kernel against the measured HBM peak: on the algorithmic bytes B and on the
                        DRAM bytes ncu measured for this build (profiles/traffic.json).
  cpu_baseline          reference-ordered CPU step on a z-slab sample: CSR SpMV of the oracle port (Trilinos is not here)
                        + the REFERENCE's own collision code (oracle/_ref, compiled from /root/reference) when present.
  parity                GPU vs oracle on a sample of the workload built by the very same code path as the timed context.
  configs               (N = 1) the other BASELINE.json configurations, each timed and gated the same way:
                        c1 TGV2D D2Q9; c3 compressible TGV3D D3Q45 f+g; c4 Riemann 2D D2Q25H f+g with walls;
                        c5 y-stretched channel D3Q45 f+g, ThermalBounceBack walls, EXACT_DIFFERENCE forcing.
  parity_multirank      (N > 1) the slab-partitioned step on a small mesh against the single-domain oracle.
Synthetic inputs (analytic fields, f = f_eq), fp64.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mdof*Q updates/s (stream+collide)"
UNIT = "MDoF*Q/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=32, help="cells per axis per GPU (32 = refinement 5)")
    ap.add_argument("--order", type=int, default=4)
    ap.add_argument("--stencil", default="D3Q19")
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--e2e-chunks", type=int, default=16, help="pieces of the DoF range nb200_step_host pipelines over (1 = sequential legs)")
    ap.add_argument("--cpu-sample-layers", type=int, default=8, help="z cell layers of the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--format", default="dict", choices=["dict", "dict-unstaged", "ell"], help="device format of the streaming matrix")
    ap.add_argument("--dof-order", default="none", choices=["cell", "none"], help="internal DoF order hint (nb200_set_dof_order)")
    ap.add_argument("--stretch", type=float, default=0.0,
                    help="side measurement: grade the mesh in y like TurbulentChannelFlow3D (y -> y - s sin(2 pi y)/(2 pi)); 0 = uniform (the bench line)")
    ap.add_argument("--row-noise", type=float, default=0.0,
                    help="worst-case matrix: every stored value gets its own relative perturbation of this size, so that no two rows share a "
                         "weight pattern (what an unstructured mesh gives); use with --dedup-tol 0")
    ap.add_argument("--jitter", type=float, default=0.0,
                    help="side measurement: move every interior mesh vertex line by a random fraction of the cell width in x, y and z "
                         "(with --dedup-tol 0 the weight-pattern pool degenerates towards one pattern per row class)")
    ap.add_argument("--numbering", default="cell", choices=["cell", "lex"],
                    help="host DoF numbering of the synthetic problem: cell-wise like deal.II (default) or lexicographic")
    ap.add_argument("--grid", default="on", choices=["on", "off"],
                    help="give the structure hint nb200_set_dof_grid (grid coordinates of the DoFs): TMA box kernels; off = staged dictionary kernels")
    ap.add_argument("--dedup-tol", type=float, default=1e-14, help="value tolerance of the dictionary format (library default)")
    ap.add_argument("--configs", default="c1,c3,c4,c5", help="other BASELINE configurations to time and gate after the headline (N = 1 only); '' = none")
    ap.add_argument("--no-gates", action="store_true", help="skip the GPU-vs-oracle parity gates")
    ap.add_argument("--no-block-partition", action="store_true", help="N > 1: skip the block-partition side line")
    return ap.parse_args()


def algorithmic_bytes_per_dof(nnz_per_dof, D, Q, with_g=False):
    """12*nnz (8 B value + 4 B index, matrix read once) + per distribution
    [8(Q-1) streamed reads + 8 rest read + 8Q writes] + 8(1+D) for rho,u [+16 for T, sensor]."""
    per_dist = 8 * (Q - 1) + 8 + 8 * Q
    return 12 * nnz_per_dof + per_dist * (2 if with_g else 1) + 8 * (1 + D) + (16 if with_g else 0)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
            t0 = time.time()
            while not self.rows and time.time() - t0 < 5.0:      # nvidia-smi needs a moment to start
                time.sleep(0.05)
            self.rows.clear()                                     # keep only samples taken under load
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy kernel)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_record(key):
    """What the committed ncu capture of this build says about the dominant kernel of workload `key` (profiles/traffic.json,
    written by tools/ncu_summary.py from the .ncu-rep): dram bytes per launch, LSU pipe and DRAM utilisation."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(key)
        except Exception:
            return None
    return None


def bind_to_gpu_numa_node(local):
    """CPU affinity (and with it first-touch placement of the pinned host buffers) on the NUMA node of this rank's GPU:
    eight ranks that pin and copy through one socket's memory controllers do not scale (r01: e2e efficiency 0.24 at N = 8)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return sorted(os.sched_getaffinity(0))
    except Exception:
        return None


# ------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------
def case_spec(key, args=None):
    """The BASELINE.json configurations as data: mesh, stencil, physics, boundary treatment."""
    Ma = 0.05
    if key == "c2":
        return dict(key="c2", dim=3, cells=[args.cells] * 3, p=args.order, stencil=args.stencil, scaling=math.sqrt(3) / Ma,
                    nu=2 * math.pi, cfl=0.4, length=[2 * math.pi] * 3, walls=[False] * 3, with_g=False, init="tgv3d",
                    name=f"TGV3D {args.stencil} BGK semi-Lagrangian p={args.order}", stretch=args.stretch, jitter=args.jitter, row_noise=args.row_noise)
    if key == "c1":
        return dict(key="c1", dim=2, cells=[8, 8], p=4, stencil="D2Q9", scaling=math.sqrt(3) / Ma, nu=1.0, cfl=0.4,
                    length=[2 * math.pi] * 2, walls=[False] * 2, with_g=False, init="tgv2d", name="TGV2D D2Q9 BGK p=4 refinement 3")
    if key == "c3":
        return dict(key="c3", dim=3, cells=[16] * 3, p=4, stencil="D3Q45", scaling=1.0, nu=0.01, cfl=0.4, length=[2 * math.pi] * 3,
                    walls=[False] * 3, with_g=True, prandtl=0.71, sutherland=True, init="tgv3d_compressible",
                    name="compressible TGV3D D3Q45 f+g quartic Pr=0.71 Sutherland p=4 refinement 4")
    if key == "c4":
        return dict(key="c4", dim=2, cells=[512, 512], p=2, stencil="D2Q25H", scaling=1.0, nu=0.001, cfl=1.0, length=[2.0, 2.0],
                    walls=[True, True], with_g=True, prandtl=None, sutherland=False, init="riemann", wall_kind="velocity",
                    name="Riemann 2D D2Q25H f+g quartic p=2 (step-11), VelocityNeqBounceBack walls on all four sides, 2x2 base cells refined 8 times")
    if key == "c5":
        return dict(key="c5", dim=3, cells=[16] * 3, p=4, stencil="D3Q45", scaling=1.0, nu=0.01, cfl=0.4, length=[2 * math.pi, 2.0, math.pi],
                    walls=[False, True, False], stretch=0.8, with_g=True, prandtl=0.7, sutherland=True, init="channel", wall_kind="thermal",
                    wall_T=0.85, force=[2e-4, 0.0, 0.0], force_type="EXACT_DIFFERENCE",
                    name="channel D3Q45 f+g quartic Pr=0.7 Sutherland p=4, y-stretched mesh (0.8), ThermalBounceBack(0.85) walls, "
                         "EXACT_DIFFERENCE forcing (step-turbulent-channel)")
    raise ValueError(key)


def case_vertices(c, cells, length):
    verts = []
    rng = np.random.default_rng(12345)
    for d in range(c["dim"]):
        t = np.arange(cells[d] + 1) / cells[d]
        if d == 1 and c.get("stretch", 0.0) > 0.0:               # TurbulentChannelFlow3D.h:116-123
            t = t - c["stretch"] * np.sin(2 * math.pi * t) / (2 * math.pi)
        if c.get("jitter", 0.0) > 0.0:
            t = t.copy()
            t[1:-1] += c["jitter"] * (rng.random(cells[d] - 1) - 0.5) / cells[d]
        verts.append(length[d] * t)
    return verts


def initial_fields(c, st, x):
    """(rho, u, T) at the support points x (physical units)."""
    from natrium_b200 import harness
    n = x.shape[0]
    if c["init"] == "tgv3d":
        rho, u = harness.taylor_green_3d(x, st.getSpeedOfSound())
        return rho, u, np.ones(n)
    if c["init"] == "tgv2d":
        rho, u = harness.taylor_green_2d(x)
        return rho, u, np.ones(n)
    if c["init"] == "tgv3d_compressible":
        rho, u = harness.taylor_green_3d(x, st.getSpeedOfSound(), compressible=True, density_numerator=0.1)
        return rho, 0.1 * u, np.ones(n)
    if c["init"] == "riemann":                                  # L/benchmarks/Riemann2D.cpp:41-106
        lo_x, lo_y = x[:, 0] <= 1.0, x[:, 1] <= 1.0
        rho = np.where(lo_x & lo_y, 0.8, np.where(~lo_x & ~lo_y, 0.5313, 1.0))
        T = np.where(lo_x & lo_y, 1.25, np.where(~lo_x & ~lo_y, 0.7532956685, 1.0))
        u = np.stack([np.where(lo_x & ~lo_y, 0.42008, 0.0), np.where(~lo_x & lo_y, 0.42008, 0.0)])
        return rho, u, T
    if c["init"] == "channel":
        H = c["length"][1]
        eta = x[:, 1] / H
        u = np.stack([0.1 * 4 * eta * (1 - eta) * (1 + 0.05 * np.sin(4 * x[:, 2])), 0.002 * np.sin(2 * x[:, 0]) * np.sin(math.pi * eta),
                      np.zeros(n)])
        return 1.0 + 0.01 * np.cos(x[:, 0]), u, 0.85 + 0.1 * 4 * eta * (1 - eta)
    raise ValueError(c["init"])


def build_product(c, cells, length, local=0, rank=0, world=1, uid=None, grid="on", fmt="dict", tol=1e-14, numbering="cell", dof_order="none",
                  blocks=None):
    """One context for case c on a mesh of `cells` (the slab of this rank when world > 1): stencil, layout, hints, matrix,
    halo, wall hits, collision, initial populations.  The timed contexts and the gate contexts all come from here."""
    from natrium_b200 import Context, harness, _capi
    from natrium_b200.stencils import Stencil
    st = Stencil(c["stencil"], c["scaling"])
    pb = harness.CartesianProblem(c["dim"], cells, c["p"], length=length, verts=case_vertices(c, cells, length))
    dt = pb.timestep(st, c["cfl"])
    ctx = Context(local, rank, world, uid)
    ctx.set_stencil(st.getDirections(), st.getWeights(), st.getScaling(), st.getSpeedOfSoundSquare())
    if blocks is not None:      # block partition (Z-curve-like): the host numbers lexicographically inside its box, order hint, staged kernels
        part = harness.BlockPartition(pb, st, dt, rank, blocks)
        numbering, grid, dof_order = "lex", "off", "cell"
    else:
        part = harness.SlabPartition(pb, st, dt, rank, world)
    with_g = bool(c.get("with_g"))
    ctx.set_layout(part.n_owned, part.n_ghost, with_g)
    fmt_code = {"dict": _capi.FORMAT_DICT, "dict-unstaged": _capi.FORMAT_DICT_UNSTAGED, "ell": _capi.FORMAT_ELL}[fmt]
    ctx.set_matrix_format(fmt_code, tol if fmt != "ell" else 0.0)
    num = harness.CellNumbering(part) if numbering == "cell" else None
    host = num if num is not None else part
    if dof_order == "cell" and num is None:
        ctx.set_dof_order(part.cell_blocked_order())
    if grid == "on" and fmt == "dict":
        ctx.set_dof_grid(*host.grid_coords(), fe_order=c["p"])
    t0 = time.perf_counter()
    hits = None
    if any(c["walls"]):
        nnz, hi, hd = harness.upload_streaming_matrix_walled(ctx, pb, part, st, dt, c["walls"], num)
        thermal = c.get("wall_kind") == "thermal"
        kinds = np.full(len(hi), _capi.WALL_THERMAL_BOUNCE_BACK if thermal else _capi.WALL_VELOCITY_NEQ_BOUNCE_BACK, dtype=np.int32)
        vals = np.full(len(hi), c.get("wall_T", 0.0) if thermal else 0.0)      # zero wall velocity: 2 w rho e.u_w / cs2 = 0
        ctx.set_wall_hits(hi, hd, kinds, vals)
        hits = (hi, hd, kinds, vals)
    else:
        nnz = harness.upload_streaming_matrix(ctx, pb, part, st, dt, num, noise=float(c.get("row_noise", 0.0)))
    t_asm = time.perf_counter() - t0
    if world > 1:
        ctx.set_halo(*host.halo_plan())
    kw = {}
    if c.get("force") is not None:
        kw = dict(force=np.array(c["force"][:c["dim"]]), force_type=getattr(_capi, c["force_type"]))
    if with_g:
        ctx.set_collision(c["nu"], dt, equilibrium=_capi.QUARTIC_EQUILIBRIUM, with_g=True, gamma=1.4, prandtl=c.get("prandtl"),
                          sutherland=bool(c.get("sutherland")), **kw)
    else:
        ctx.set_collision(c["nu"], dt, **kw)
    x = host.owned_points()
    rho, u, T = initial_fields(c, st, x)
    if with_g:
        f0, g0 = harness.quartic_equilibrium_distributions(st, rho, u, T, 1.4)
        ctx.upload_populations(1, g0)
    else:
        f0, g0 = harness.equilibrium_distributions(st, rho, u), None
    ctx.upload_populations(0, f0)
    return dict(ctx=ctx, st=st, pb=pb, part=part, host=host, num=num, dt=dt, nnz=nnz, t_asm=t_asm, f0=f0, g0=g0, hits=hits, n=part.n_owned)


def kernel_name(ctx, c):
    st = c["stencil"]
    D, Q = (2 if st.startswith("D2") else 3), int("".join(ch for ch in st.split("Q")[1] if ch.isdigit()))
    gi = ctx.grid_info()
    staged = ctx.matrix_format_info().get("staged")
    fused = Q <= 25 and c.get("force") is None and not any(c["walls"])
    fam = "grid" if gi["in_use"] else ("staged" if staged else "rows")
    if fused:
        return f"k_stream_collide_{'fg' if c.get('with_g') else 'f'}_{fam}<{D},{Q}>"
    return (f"k_stream_{fam}<{D},{Q},{2 if c.get('with_g') and not any(c['walls']) else 1}> (+ k_wall_hits)" if any(c["walls"]) else
            f"k_stream_{fam}<{D},{Q},{2 if c.get('with_g') else 1}>") + f" + k_collide_{'fg' if c.get('with_g') else 'f'}<{D},{Q}>"


# ------------------------------------------------------------------------------------------
# checker legs (the only places that touch oracle/): parity gates and the CPU baseline
# ------------------------------------------------------------------------------------------
def oracle_stepper_for(c, st_name_scaling, blocks, n, dt, hits):
    """Reference-ordered CPU step for case c on the given blocks (dict (bi,bj) -> scipy CSR): returns step(f, g) -> status."""
    from oracle import cpu
    ost = cpu.Stencil(*st_name_scaling)
    with_g = bool(c.get("with_g"))
    if hits is None and c.get("force") is None:
        stepper = cpu.ReferenceOrderStepper(ost, blocks, n, c["nu"], dt, equilibrium=1 if with_g else 0, with_g=with_g, gamma=1.4,
                                            prandtl=c.get("prandtl"), sutherland=bool(c.get("sutherland")))
        return lambda f, g: stepper.step(f, g)
    hi, hd, kinds, vals = hits if hits is not None else (np.zeros(0, np.int32),) * 3 + (np.zeros(0),)

    def step(f, g):      # stream(f) -> wall hits -> stream(g) -> collide (CompressibleCFDSolver.h:181-314)
        f[...] = cpu.stream(blocks, f)
        if len(hi) and cpu.apply_wall_hits(ost, f, g, hi, hd, kinds, vals) != 0:
            return 1
        if with_g:
            g[...] = cpu.stream(blocks, g)
            if c.get("force") is not None:
                rc = cpu.collide_bgk_fg_forced(ost, f, g, c["nu"], dt, np.array(c["force"][:c["dim"]]), c["force_type"], gamma=1.4,
                                               prandtl=c.get("prandtl"), sutherland=bool(c.get("sutherland")))[-1]
            else:
                rc = cpu.collide_bgk_fg(ost, f, g, c["nu"], dt, equilibrium=1, gamma=1.4, prandtl=c.get("prandtl"), sutherland=bool(c.get("sutherland")))[-1]
        else:
            rc = cpu.collide_bgk(ost, f, c["nu"], dt)[-1]
        return rc
    return step


def harness_blocks(c, B):
    """The matrix of the product context B once more as scipy CSR blocks in the HOST numbering (input of the oracle stepper)."""
    import scipy.sparse as sp
    from natrium_b200 import harness
    pb, part, st, dt, num, n = B["pb"], B["part"], B["st"], B["dt"], B["num"], B["n"]
    blocks = {}
    for a in range(1, st.getQ()):
        if any(c["walls"]):
            bl, _ = harness.assemble_direction_walled(pb, part, st, dt, a, c["walls"])
        else:
            rp0, col0, val0 = harness.assemble_direction(pb, part, st, dt, a)
            if c.get("row_noise", 0.0) > 0.0:
                val0 = harness.row_noise(val0, a, float(c["row_noise"]))
            bl = {(a - 1, a - 1): (rp0, col0, val0)}
        for k, (rp, col, val) in bl.items():
            if num is not None:
                rp, col, val = num.renumber_csr(rp, col, val)
            if len(val):
                blocks[k] = sp.csr_matrix((val, col, rp), shape=(n, n))
    return blocks


def parity_gate(c, cells, length, args, steps=10):
    """GPU vs oracle over `steps` steps from the identical state, per-step max relative error of f (and g), on a context
    built by build_product with the options of the timed run (format, tolerance, numbering, hints) -- SURVEY 8d."""
    B = build_product(c, cells, length, local=int(os.environ.get("LOCAL_RANK", "0")), grid=args.grid, fmt=args.format, tol=args.dedup_tol,
                      numbering=args.numbering, dof_order=args.dof_order)
    ctx, n, Q = B["ctx"], B["n"], B["st"].getQ()
    try:
        blocks = harness_blocks(c, B)
        ones = np.ones((Q, n))
        with_g = bool(c.get("with_g"))
        step = oracle_stepper_for(c, (c["stencil"], c["scaling"]), blocks, n, B["dt"], B["hits"])
        f, g = B["f0"].copy(), (B["g0"].copy() if with_g else None)
        worst = 0.0
        for s in range(steps):
            ctx.step(1)
            ctx.synchronize()
            if step(f, g) != 0:
                raise RuntimeError("oracle step failed (density)")
            e = float(np.max(np.abs(ctx.download_populations(0) - f) / np.maximum(np.abs(f), 1e-300)))
            if with_g:
                e = max(e, float(np.max(np.abs(ctx.download_populations(1) - g) / np.maximum(np.abs(g), 1e-300))))
            worst = max(worst, e)
            ctx.upload_populations(0, f)                      # identical state for the next step
            if with_g:
                ctx.upload_populations(1, g)
        # row sums: M 1 = 1 (SemiLagrangian_test.cpp:519-596), bounce blocks included; the hit list is taken off for it
        # (a ThermalBounceBack hit re-equilibrates its DoF, which is not part of the matrix)
        if B["hits"] is not None:
            ctx.set_wall_hits(np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0))
        ctx.upload_populations(0, ones)
        ctx.stream(0)
        row_sum_err = float(np.max(np.abs(ctx.download_populations(0)[1:] - 1.0)))
        gi = ctx.grid_info()
        return {"steps": steps, "max_rel_err": worst, "row_sum_err": row_sum_err, "tolerance": 1e-12,
                # (a matrix perturbed by --row-noise has no unit row sums)
                "ok": bool(worst <= 1e-12 and (row_sum_err <= 1e-12 or c.get("row_noise", 0.0) > 0.0)),
                "mesh": "x".join(str(v) for v in cells) + " cells", "n_dofs": n,
                "kernels": "grid (TMA boxes)" if gi["in_use"] else ("staged" if ctx.matrix_format_info().get("staged") else "rows"),
                "built_like_timed_context": {"format": args.format, "dedup_tol": args.dedup_tol, "numbering": args.numbering, "grid_hint": args.grid,
                                             "dof_order": args.dof_order}}
    finally:
        ctx.close()


def cpu_reference_run(args, steps, warmup, layers, target_s=12.0):
    """Reference-ordered CPU step on a z-slab sample of the headline workload, all host threads: full-vector copy + one CSR
    SpMV per block (oracle port with OpenMP: Trilinos is not in this image) + collideAll -- the REFERENCE's own
    selectCollision (oracle/_ref) on per-thread row ranges when the library is there, else the oracle port."""
    from natrium_b200 import harness
    from natrium_b200.stencils import Stencil
    import scipy.sparse as sp
    from oracle import cpu
    c = case_spec("c2", args)
    st = Stencil(c["stencil"], c["scaling"])
    full = harness.CartesianProblem(3, c["cells"], c["p"], length=c["length"], verts=case_vertices(c, c["cells"], c["length"]))
    dt = full.timestep(st, 0.4)
    cells = [args.cells, args.cells, min(layers, args.cells)]
    L = [2 * math.pi, 2 * math.pi, 2 * math.pi * cells[2] / args.cells]
    pb = harness.CartesianProblem(3, cells, args.order, length=L, verts=case_vertices(c, cells, L))
    part = harness.SlabPartition(pb, st, dt)
    n = part.n_owned
    blocks = {}
    for a in range(1, st.getQ()):
        rp, col, val = harness.assemble_direction(pb, part, st, dt, a)
        blocks[(a - 1, a - 1)] = sp.csr_matrix((val, col, rp), shape=(n, n))
    ost = cpu.Stencil(c["stencil"], c["scaling"])
    rho, u = harness.taylor_green_3d(part.owned_points(), st.getSpeedOfSound())
    f = harness.equilibrium_distributions(st, rho, u)
    stepper = cpu.ReferenceOrderStepper(ost, blocks, n, c["nu"], dt)
    try:          # torchrun exports OMP_NUM_THREADS=1 to its workers, which would turn the CPU arm into a single-core run
        n_thr = len(os.sched_getaffinity(0))
    except AttributeError:
        n_thr = os.cpu_count() or 1
    cpu.set_num_threads(n_thr)
    kind, step = "port", (lambda: stepper.step(f))
    try:
        from oracle import ref
        if ref.available():
            from concurrent.futures import ThreadPoolExecutor
            pool = ThreadPoolExecutor(n_thr)
            bounds = [n * t // n_thr for t in range(n_thr + 1)]
            scratch = [(np.zeros(bounds[t + 1] - bounds[t]), np.zeros((3, bounds[t + 1] - bounds[t]))) for t in range(n_thr)]

            def collide_range(t):
                a, b = bounds[t], bounds[t + 1]
                if b > a:
                    rc = ref.select_collision_range(c["stencil"], c["scaling"], f, a, b, c["nu"], dt, scratch[t][0], scratch[t][1])
                    if rc != 0:
                        raise RuntimeError(f"reference collide failed ({rc})")

            def step_ref():
                f[...] = cpu.stream(blocks, f)                      # f_tmp = f; f.FStream = M f_tmp.FStream (CFDSolver.cpp:671-672)
                list(pool.map(collide_range, range(n_thr)))         # selectCollision -> collideAll (one MPI rank per thread)
            kind, step = "port+ref-collide", step_ref
    except Exception:
        pass
    for _ in range(warmup):
        step()
    if steps <= 0:                       # calibrate: about `target_s` seconds of CPU work
        t0 = time.perf_counter()
        step()
        steps = int(max(3, min(2000, target_s / max(1e-4, time.perf_counter() - t0))))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    el = time.perf_counter() - t0
    assert np.isfinite(f).all()
    val = n * st.getQ() * steps / el / 1e6
    port_val = None
    if kind != "port":       # for orientation: the same step with the oracle's own collide (which hoists what the reference
        t0 = time.perf_counter()        # recomputes per DoF: the Hermite tensors H3 / H4, CollisionOperator.h:68-69)
        k = 0
        while time.perf_counter() - t0 < 2.0 or k < 3:
            stepper.step(f)
            k += 1
        port_val = n * st.getQ() * k / (time.perf_counter() - t0) / 1e6
    return dict(value=val, unit=UNIT, cores=n_thr, kind=kind, port_only_value=port_val,
                sample=f"{cells[0]}x{cells[1]}x{cells[2]} cells of the {args.cells}^3 workload ({n} DoFs, {stepper.b.nnz} nnz), {steps} reference-ordered "
                       f"steps (copy + {st.getQ()-1} CSR SpMV: oracle C port with OpenMP; collide: "
                       f"{'the reference`s own selectCollision compiled from /root/reference (oracle/_ref), one row range per thread' if kind != 'port' else 'oracle C port'}), "
                       f"{el:.1f} s"), el / steps * 1e3, steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 400))          # each step is a bounded sample: keeps the whole run within minutes
    warmup = max(0, args.warmup)
    base, ms, steps = cpu_reference_run(args, steps, warmup, args.cpu_sample_layers)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, args.gpus),
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "NATriuM as a whole cannot be built here (deal.II/Trilinos/p4est/Boost/MPI absent): the SpMV is the oracle's "
                    "restatement of the Epetra loop, the collision is the reference's own code (oracle/_ref) when kind says so; "
                    "steps are capped at 400 bounded-sample steps"}
    print(json.dumps(line), flush=True)


def workload_config(args, n_gpus):
    nd = args.cells * args.order + 1
    return {"workload": f"TGV3D {args.stencil} BGK semi-Lagrangian p={args.order} {args.cells}^3 cells/GPU "
                        f"({nd}^3 DoFs/GPU) x {n_gpus} GPU slab(s) along z",
            "cells_per_gpu": args.cells ** 3, "fe_order": args.order, "stencil": args.stencil, "collision": "BGK_STANDARD",
            "cfl": 0.4, "mach": 0.05, "y_stretch": args.stretch, "vertex_jitter": args.jitter, "row_noise": args.row_noise,
            "parallelism": f"slab x{n_gpus} (NCCL ghost exchange)" if n_gpus > 1 else "single GPU",
            "l2_policy": "inputs_exceed_l2 (populations + matrix tables streamed per step > 126 MB L2)"}


# ------------------------------------------------------------------------------------------
def time_steps(ctx, steps, warmup, barrier):
    ctx.step(warmup)
    barrier()
    l0 = ctx.kernel_launches()
    ctx.timer_start()
    ctx.step(steps)
    ms = ctx.timer_stop()
    barrier()
    return ms, ctx.kernel_launches() - l0


def run_side_config(key, args, local):
    """One of the other BASELINE configurations on one GPU: device-resident ms/step, roofline on B, e2e for f+g, gate."""
    import torch
    c = case_spec(key, args)
    B = build_product(c, c["cells"], c["length"], local=local, grid=args.grid, fmt=args.format, tol=args.dedup_tol,
                      numbering=args.numbering, dof_order=args.dof_order)
    ctx, st, n = B["ctx"], B["st"], B["n"]
    Q, D = st.getQ(), st.getD()
    with_g = bool(c.get("with_g"))
    out = {"workload": c["name"] + f", {'x'.join(str(v) for v in c['cells'])} cells, {n} DoFs", "n_dofs": n, "nnz": B["nnz"],
           "wall_hits": 0 if B["hits"] is None else int(len(B["hits"][0]))}
    try:
        def barrier():
            ctx.synchronize()
            torch.cuda.synchronize()
        ctx.collide()
        est = 3e-9 * B["nnz"] * (2 if with_g else 1) / 1e3 + 0.05            # rough ms/step, to size the timed region
        steps = int(max(20, min(args.steps, 300, 1500.0 / est)))        # side lines: a few hundred steps are plenty (and keep the run short)
        ms, launches = time_steps(ctx, steps, max(3, args.warmup), barrier)
        cons = ctx.conserved()
        peak, peak_src = measured_peak()
        bpd = algorithmic_bytes_per_dof(B["nnz"] / n, D, Q, with_g)
        kms = ms / steps
        rec = ncu_record(f"{key}_{args.grid}") or {}
        out.update({"ms_per_step": kms, "steps": steps, "value": n * Q * steps / (ms * 1e-3) / 1e6, "unit": UNIT,
                    "gpu_launches_per_step": launches / steps,
                    # walled cases whose hit values are all zero (resting walls) keep the fused kernel: one launch per step
                    "kernels": kernel_name(ctx, dict(c, walls=[False] * c["dim"]) if launches == steps and any(c["walls"]) else c),
                    "roofline": {"bound": "hbm", "achieved": bpd * n / (kms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                 "frac": bpd * n / (kms * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_dof": bpd,
                                 "traffic": rec.get("dram_bytes_per_step"),
                                 "frac_on_dram_bytes": (rec["dram_bytes_per_step"] / (kms * 1e-3) / 1e9 / peak) if rec.get("dram_bytes_per_step") else None,
                                 "peak_source": peak_src},
                    "matrix_format": ctx.matrix_format_info(), "grid": ctx.grid_info(), "conserved_finite": bool(np.isfinite(cons).all())})
        # end to end through host buffers
        e2e_steps = max(2, min(args.e2e_steps, 10))
        hf = [torch.empty((Q, n), dtype=torch.float64, pin_memory=True) for _ in range(2)]
        hf[0].numpy()[...] = ctx.download_populations(0)
        hmom = torch.empty((2 + D, n), dtype=torch.float64, pin_memory=True)
        if with_g:
            hg = [torch.empty((Q, n), dtype=torch.float64, pin_memory=True) for _ in range(2)]
            hg[0].numpy()[...] = ctx.download_populations(1)

            def e2e_step(i):
                ctx.step_host_fg(hf[i & 1].data_ptr(), hg[i & 1].data_ptr(), hf[(i + 1) & 1].data_ptr(), hg[(i + 1) & 1].data_ptr(),
                                 hmom.data_ptr(), hmom.data_ptr() + 8 * n, hmom.data_ptr() + 8 * n * (1 + D))
            h2d, d2h = 2 * Q * n * 8, 2 * Q * n * 8 + (2 + D) * n * 8
        else:
            def e2e_step(i):
                ctx.step_host(hf[i & 1].data_ptr(), hf[(i + 1) & 1].data_ptr(), hmom.data_ptr(), hmom.data_ptr() + 8 * n, args.e2e_chunks)
            h2d, d2h = Q * n * 8, Q * n * 8 + (1 + D) * n * 8
        e2e_step(0); e2e_step(1)
        barrier()
        ctx.timer_start()
        for i in range(e2e_steps):
            e2e_step(i)
        ms_e = ctx.timer_stop()
        barrier()
        out["e2e"] = {"value": n * Q * e2e_steps / (ms_e * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms_e / e2e_steps, "steps": e2e_steps,
                      "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                      "api": "nb200_step_host_fg" if with_g else "nb200_step_host"}
    finally:
        ctx.close()
    if not args.no_gates:
        gate_cells = {"c1": [8, 8], "c3": [6, 6, 2], "c4": [48, 48], "c5": [6, 6, 2]}[key]
        gl = [c["length"][d] * gate_cells[d] / c["cells"][d] for d in range(c["dim"])]
        if any(c["walls"]):
            gl = [gl[d] if not c["walls"][d] else c["length"][d] for d in range(c["dim"])]     # walled axes keep the full width
        try:
            out["parity"] = parity_gate(c, gate_cells, gl, args, steps=5 if Q == 45 else 10)
        except Exception as ex:
            out["parity"] = {"ok": False, "error": str(ex)}
    return out


BLOCKS_OF_WORLD = {2: [2, 1, 1], 4: [2, 2, 1], 8: [2, 2, 2]}


def multirank_parity(args, rank, world, local, uid, blocks=None):
    """N > 1: the slab-partitioned (or, blocks given, block-partitioned) step (ghost exchange over NCCL, interior / boundary
    split) on a small mesh against the single-domain oracle, gathered on rank 0."""
    import torch
    import torch.distributed as dist
    from natrium_b200 import harness
    c = case_spec("c2", args)
    cells = [4, 4, 2 * world] if blocks is None else [2 * b for b in blocks]
    L = [2 * math.pi, 2 * math.pi, 2 * math.pi]
    from natrium_b200 import Context
    box = [Context.unique_id() if rank == 0 else None]          # a communicator of its own: an NCCL id is good for one
    dist.broadcast_object_list(box, src=0)
    uid = box[0]
    B = build_product(c, cells, L, local=local, rank=rank, world=world, uid=uid, grid=args.grid, fmt=args.format, tol=args.dedup_tol,
                      numbering=args.numbering, dof_order=args.dof_order, blocks=blocks)
    ctx = B["ctx"]
    steps = 5
    ctx.step(steps)
    ctx.synchronize()
    got = ctx.download_populations(0)
    gids = B["part"].owned_global_ids()
    if B["num"] is not None:
        gids = gids[B["num"].order]
    ctx.close()
    parts = [None] * world if rank == 0 else None
    dist.gather_object((gids, got), parts, dst=0)
    if rank != 0:
        return None
    one = build_single_domain_oracle(c, cells, L, steps)
    full = np.zeros_like(one)
    for g, a in parts:
        full[:, g] = a
    err = float(np.max(np.abs(full - one) / np.maximum(np.abs(one), 1e-300)))
    return {"mesh": "x".join(str(v) for v in cells) + " cells over " + (str(world) + " slabs" if blocks is None else "x".join(str(b) for b in blocks) + " blocks"),
            "steps": steps, "max_rel_err": err, "tolerance": 1e-11, "ok": bool(err <= 1e-11)}


def block_partition_line(args, rank, world, local):
    """N in {2, 4, 8}: the headline problem family on a block partition (2x1x1 / 2x2x1 / 2x2x2 blocks of 16^3 cells: up to 6
    neighbours, edge ghosts, a cut across the x-fastest numbering) -- what the reference's p4est Z-curve partition looks like
    (L/advection/SemiLagrangian.cpp:185-193) -- timed device-resident next to the slab headline, with its own parity gate."""
    import torch
    import torch.distributed as dist
    from natrium_b200 import Context
    blocks = BLOCKS_OF_WORLD[world]
    c = case_spec("c2", args)
    per = 16
    cells = [per * b for b in blocks]
    L = [2 * math.pi * b for b in blocks]
    box = [Context.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    B = build_product(c, cells, L, local=local, rank=rank, world=world, uid=box[0], fmt=args.format, tol=args.dedup_tol, blocks=blocks)
    ctx, st, pb, part = B["ctx"], B["st"], B["pb"], B["part"]
    n_nbr = len(part.halo_plan()[0])

    def barrier():
        ctx.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
    ctx.collide()
    steps = max(20, min(args.steps, 500))
    ms, _ = time_steps(ctx, steps, max(3, args.warmup), barrier)
    t = torch.tensor([ms, float(n_nbr)], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, max_nbr = float(t[0].item()), int(t[1].item())
    cons = ctx.conserved()
    info = ctx.matrix_format_info()
    ctx.close()
    parity = multirank_parity(args, rank, world, local, None, blocks=blocks)
    if rank != 0:
        return None
    return {"workload": f"TGV3D D3Q19 BGK p={c['p']}, {'x'.join(str(b) for b in blocks)} blocks of {per}^3 cells ({pb.N} DoFs)", "blocks": blocks,
            "max_neighbours": max_nbr, "ms_per_step": ms / steps, "steps": steps, "value": pb.N * st.getQ() * steps / (ms * 1e-3) / 1e6, "unit": UNIT,
            "kernels": "staged dictionary kernels (no grid hint), interior / boundary split", "staged": info.get("staged"),
            "conserved_finite": bool(np.isfinite(cons).all()), "parity": parity}


def build_single_domain_oracle(c, cells, L, steps):
    """`steps` reference-ordered CPU steps of case c on the whole (unpartitioned) mesh, lexicographic numbering."""
    import scipy.sparse as sp
    from natrium_b200 import harness
    from natrium_b200.stencils import Stencil
    st = Stencil(c["stencil"], c["scaling"])
    pb = harness.CartesianProblem(c["dim"], cells, c["p"], length=L, verts=case_vertices(c, cells, L))
    dt = pb.timestep(st, c["cfl"])
    part = harness.SlabPartition(pb, st, dt)
    n = part.n_owned
    blocks = {}
    for a in range(1, st.getQ()):
        rp, col, val = harness.assemble_direction(pb, part, st, dt, a)
        blocks[(a - 1, a - 1)] = sp.csr_matrix((val, col, rp), shape=(n, n))
    rho, u, T = initial_fields(c, st, part.owned_points())
    f = harness.equilibrium_distributions(st, rho, u)
    step = oracle_stepper_for(c, (c["stencil"], c["scaling"]), blocks, n, dt, None)
    for _ in range(steps):
        assert step(f, None) == 0
    return f


def run_ours(args):
    import torch
    import torch.distributed as dist
    from natrium_b200 import Context

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local)
    affinity = bind_to_gpu_numa_node(local) if world > 1 else None
    uid = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        box = [Context.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]

    c = case_spec("c2", args)
    cells = [args.cells, args.cells, args.cells * world]
    L3 = [2 * math.pi, 2 * math.pi, 2 * math.pi * world]
    B = build_product(c, cells, L3, local=local, rank=rank, world=world, uid=uid, grid=args.grid, fmt=args.format, tol=args.dedup_tol,
                      numbering=args.numbering, dof_order=args.dof_order)
    ctx, st, pb, part, n, nnz = B["ctx"], B["st"], B["pb"], B["part"], B["n"], B["nnz"]
    Q, D = st.getQ(), st.getD()
    ctx.collide()               # run(): collide once before the loop

    def barrier():
        ctx.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing
    ctx.step(args.warmup)
    barrier()
    launches0 = ctx.kernel_launches()
    sampler = ClockSampler(local)
    sampler.start()
    ctx.timer_start()
    ctx.step(args.steps)
    ms = ctx.timer_stop()
    barrier()
    launches = ctx.kernel_launches() - launches0
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if ms < 400.0:
        # a short timed region ends before nvidia-smi (50 ms period) delivers samples: keep the same kernel running,
        # untimed, for ~0.4 s so that the clocks under this load are seen; the step count is derived from the
        # all-reduced time, i.e. identical on every rank (each step contains a collective)
        ctx.step(int(math.ceil(400.0 / max(ms / args.steps, 1e-3))))
        barrier()
    clocks = sampler.stop()
    n_global = pb.N
    value = n_global * Q * args.steps / (ms * 1e-3) / 1e6
    cons = ctx.conserved()
    assert np.isfinite(cons).all()

    # ---- end to end: host buffers in, host buffers out, every step (nb200_step_host: the call a host-resident
    # DistributionFunctions makes; H2D of f, fused step, D2H of f, rho, u -- pipelined over chunks of the DoF range)
    hf_a = torch.empty((Q, n), dtype=torch.float64, pin_memory=True)
    hf_b = torch.empty((Q, n), dtype=torch.float64, pin_memory=True)
    hf_a.numpy()[...] = ctx.download_populations(0)
    hmom = torch.empty((1 + D, n), dtype=torch.float64, pin_memory=True)      # rho, u: pinned like f
    e2e_steps = max(1, args.e2e_steps)
    bufs = [hf_a, hf_b]
    rho_ptr, u_ptr = hmom.data_ptr(), hmom.data_ptr() + 8 * n

    e2e_chunks = args.e2e_chunks

    def e2e_step(i):
        ctx.step_host(bufs[i & 1].data_ptr(), bufs[(i + 1) & 1].data_ptr(), rho_ptr, u_ptr, e2e_chunks)

    def e2e_time(n_steps):
        e2e_step(0)
        e2e_step(1)
        barrier()
        ctx.timer_start()
        for i in range(n_steps):
            e2e_step(i)
        ms_ = ctx.timer_stop()
        barrier()
        if world > 1:
            t_ = torch.tensor([ms_], device="cuda", dtype=torch.float64)
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
            ms_ = float(t_.item())
        return ms_

    e2e_probe = None
    if world > 1 and args.e2e_chunks > 1:
        # n_chunks is the caller's knob (include/natrium_b200.h): with several GPUs under one host the duplex chunk pipeline and
        # the sequential legs load the PCIe root / host memory differently, so a short probe of both picks the setting that is
        # then timed (every rank sees the same all-reduced times and takes the same branch)
        e2e_probe = {}
        for ch in (args.e2e_chunks, 1):
            e2e_chunks = ch
            e2e_probe[str(ch)] = e2e_time(2) / 2
        e2e_chunks = min(e2e_probe, key=e2e_probe.get)
        e2e_chunks = int(e2e_chunks)
    ms_e2e = e2e_time(e2e_steps)
    assert np.isfinite(hmom.numpy()).all() and np.isfinite(bufs[e2e_steps & 1].numpy()).all()
    e2e_val = n_global * Q * e2e_steps / (ms_e2e * 1e-3) / 1e6
    h2d = Q * n * 8
    d2h = Q * n * 8 + (1 + D) * n * 8

    # ---- roofline of the dominant kernel (the fused stream+collide kernel: one launch per step per GPU)
    peak, peak_src = measured_peak()
    bpd = algorithmic_bytes_per_dof(nnz / n, D, Q, False)
    alg_bytes = bpd * n
    kern_ms = ms / args.steps          # N=1: the step IS the kernel launch; N>1 includes the halo exchange
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    info = ctx.matrix_info()
    gi = ctx.grid_info()
    rec = ncu_record(f"{args.stencil}_p{args.order}_{args.cells}_{'grid' if gi['in_use'] else args.format}") or {}
    traffic = rec.get("dram_bytes_per_launch") if (args.stretch == 0.0 and args.jitter == 0.0 and args.row_noise == 0.0 and args.dedup_tol == 1e-14) else None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "frac_on_dram_bytes": (traffic / (kern_ms * 1e-3) / 1e9 / peak) if traffic else None,
                "lsu_pipe_frac_ncu": rec.get("lsu_wavefronts_pct_of_peak"), "dram_frac_ncu": rec.get("dram_throughput_pct_of_peak"),
                "ncu_capture": rec.get("capture"),
                "peak_source": peak_src, "algorithmic_bytes_per_dof": bpd, "algorithmic_bytes_per_launch": alg_bytes,
                "kernel": kernel_name(ctx, c), "kernel_ms": kern_ms, "frac_of_nominal_8000": achieved / 8000.0,
                "device_format_bytes": info["device_bytes"], "nnz": nnz,
                "note": "frac is defined on the algorithmic bytes B = 12*nnz + populations (SURVEY 8d); the dictionary format stores every distinct "
                        "weight pattern once and the grid kernels fetch support values as TMA boxes, so the DRAM bytes (traffic) are several times "
                        "smaller than B and frac > 1; the kernel's limiter is the L1/shared-memory data pipe (lsu_pipe_frac_ncu), not HBM",
                "matrix_format": ctx.matrix_format_info(), "grid": gi, "dof_order": args.dof_order, "host_numbering": args.numbering}

    ctx.close()
    parity = None
    if rank == 0 and world == 1 and not args.no_gates:
        try:          # a sample of the workload (same x-y extent, 4 cell layers in z), built by the same code path
            parity = parity_gate(c, [args.cells, args.cells, 4], [2 * math.pi, 2 * math.pi, 2 * math.pi * 4 / args.cells], args)
            if not parity["ok"]:
                print(f"PARITY GATE FAILED: {parity}", file=sys.stderr, flush=True)
        except Exception as ex:
            parity = {"ok": False, "error": str(ex)}
    parity_mr = None
    if world > 1 and not args.no_gates:
        try:
            parity_mr = multirank_parity(args, rank, world, local, uid)
        except Exception as ex:
            parity_mr = {"ok": False, "error": str(ex)}
    block_line = None
    if world in BLOCKS_OF_WORLD and not args.no_block_partition:
        try:
            block_line = block_partition_line(args, rank, world, local)
        except Exception as ex:       # every rank takes the same path up to its own failure; the line survives
            block_line = {"error": str(ex)}

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu_base, _, _ = cpu_reference_run(args, 0, 1, args.cpu_sample_layers)
        except Exception as ex:    # the baseline is a reported number, never a reason to lose the bench line
            cpu_base = {"value": None, "unit": UNIT, "cores": None, "kind": "port", "sample": f"failed: {ex}"}

    configs = None
    if rank == 0 and world == 1 and args.configs.strip() and args.stencil == "D3Q19":
        configs = {}
        for key in [k.strip() for k in args.configs.split(",") if k.strip()]:
            try:
                configs[key] = run_side_config(key, args, local)
            except Exception as ex:
                configs[key] = {"error": str(ex)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
                "clocks": clocks,
                "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                        "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps, "chunks": e2e_chunks, "chunks_probe_ms_per_step": e2e_probe,
                        "api": "nb200_step_host (pinned host buffers in and out every step)", "cpu_affinity_of_rank0": affinity},
                "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu_base, "parity": parity,
                "parity_multirank": parity_mr, "partition_block": block_line, "configs": configs,
                "n_dofs_global": n_global, "matrix_assembly_upload_s": B["t_asm"],
                "conserved": [float(x) for x in cons]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
