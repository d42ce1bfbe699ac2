#!/usr/bin/env python
"""bench.py -- throughput of the fused stream+collide step (BASELINE.json metric).

  python bench.py --gpus 1 --steps K --warmup W            # our arm (CUDA, through the C ABI)
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   # weak scaling, one rank per GPU
  python bench.py --impl reference ...                     # the reference's CPU path (oracle port) on host cores

Workload (config.workload): BASELINE.json configs[1] -- Taylor-Green vortex 3D, D3Q19 BGK, semi-Lagrangian
streaming, FE order 4, refinement 5 (32^3 cells, 129^3 = 2 146 689 DoFs, 7.08e8 matrix non-zeros) per GPU;
N GPUs hold N such slabs stacked along z (weak scaling), ghost planes exchanged over NCCL each step.
Synthetic inputs (analytic TGV fields, f = f_eq), fp64.

One JSON line on stdout (rank 0).  `value` = all-rank DoF*Q updates per second with everything resident in
HBM; `e2e` = the same step driven with HOST buffers (pinned H2D of f, step, D2H of f + rho,u) through the C ABI.
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
    ap.add_argument("--numbering", default="cell", choices=["cell", "lex"],
                    help="host DoF numbering of the synthetic problem: cell-wise like deal.II (default) or lexicographic")
    ap.add_argument("--grid", default="on", choices=["on", "off"],
                    help="give the structure hint nb200_set_dof_grid (grid coordinates of the DoFs): TMA box kernels; off = staged dictionary kernels")
    ap.add_argument("--dedup-tol", type=float, default=1e-14, help="value tolerance of the dictionary format (library default)")
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


def ncu_traffic(workload_key):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if one exists."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(workload_key)
        except Exception:
            return None
    return None


# ------------------------------------------------------------------------------------------
def cpu_reference_run(args, steps, warmup, layers, target_s=12.0, gpu_parity=False):
    """Reference-ordered CPU step (oracle port, OpenMP over rows/DoFs) on a z-slab sample of the workload."""
    from natrium_b200 import harness
    from natrium_b200.stencils import Stencil
    import scipy.sparse as sp
    from oracle import cpu

    Ma = 0.05
    st = Stencil(args.stencil, math.sqrt(3) / Ma)
    full = harness.CartesianProblem(3, args.cells, args.order)
    dt = full.timestep(st, 0.4)
    cells = [args.cells, args.cells, min(layers, args.cells)]
    L = [2 * math.pi, 2 * math.pi, 2 * math.pi * cells[2] / args.cells]
    pb = harness.CartesianProblem(3, cells, args.order, length=L)
    part = harness.SlabPartition(pb, st, dt)
    n = part.n_owned
    blocks = {}
    for a in range(1, st.getQ()):
        rp, col, val = harness.assemble_direction(pb, part, st, dt, a)
        blocks[(a - 1, a - 1)] = sp.csr_matrix((val, col, rp), shape=(n, n))
    ost = cpu.Stencil(args.stencil, math.sqrt(3) / Ma)
    x = part.owned_points()
    rho, u = harness.taylor_green_3d(x, st.getSpeedOfSound())
    f = harness.equilibrium_distributions(st, rho, u)
    stepper = cpu.ReferenceOrderStepper(ost, blocks, n, 2 * math.pi, dt)
    # all the host threads there are: torchrun exports OMP_NUM_THREADS=1 to its workers, which would turn the CPU arm
    # into a single-core run
    try:
        cpu.set_num_threads(len(os.sched_getaffinity(0)))
    except AttributeError:
        cpu.set_num_threads(os.cpu_count() or 1)
    parity = None
    if gpu_parity:
        # parity gate that accompanies the throughput number: the GPU path on this very sample, 10 steps from
        # the identical state, against the oracle -- max relative error per population value
        from natrium_b200 import Context
        ctx = Context(int(os.environ.get("LOCAL_RANK", "0")))
        ctx.set_stencil(st.getDirections(), st.getWeights(), st.getScaling(), st.getSpeedOfSoundSquare())
        ctx.set_layout(part.n_owned, part.n_ghost, False)
        ctx.set_dof_order(part.cell_blocked_order())
        for (bi, bj), m in blocks.items():
            ctx.upload_block_csr(bi, bj, m.indptr, m.indices, m.data)
        ctx.finalize_matrix()
        ones = np.ones((st.getQ(), n))
        ctx.upload_populations(0, ones)
        ctx.stream(0)
        row_sum_err = float(np.max(np.abs(ctx.download_populations(0) - 1.0)))      # M 1 = 1 (SemiLagrangian_test.cpp:519-596)
        ctx.set_collision(2 * math.pi, dt)
        ctx.upload_populations(0, f)
        ctx.step(10)
        ctx.synchronize()
        got = ctx.download_populations(0)
        fr = f.copy()
        for _ in range(10):
            stepper.step(fr)
        parity = {"steps": 10, "max_rel_err": float(np.max(np.abs(got - fr) / np.abs(fr))), "row_sum_err": row_sum_err,
                  "staged": bool(ctx.matrix_format_info()["staged"]), "sample": f"{cells[0]}x{cells[1]}x{cells[2]} cells"}
        ctx.close()
    for _ in range(warmup):
        stepper.step(f)
    if steps <= 0:                       # calibrate: about `target_s` seconds of CPU work
        t0 = time.perf_counter()
        stepper.step(f)
        steps = int(max(3, min(2000, target_s / max(1e-4, time.perf_counter() - t0))))
    t0 = time.perf_counter()
    for _ in range(steps):
        stepper.step(f)
    el = time.perf_counter() - t0
    assert np.isfinite(f).all()
    val = n * st.getQ() * steps / el / 1e6
    return dict(value=val, unit=UNIT, cores=cpu.num_threads(), kind="port",
                sample=f"{cells[0]}x{cells[1]}x{cells[2]} cells of the {args.cells}^3 workload ({n} DoFs, "
                       f"{stepper.b.nnz} nnz), {steps} reference-ordered steps (copy + {st.getQ()-1} CSR SpMV + collide), "
                       f"oracle C port with OpenMP, {el:.1f} s", parity_vs_gpu=parity), el / steps * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 400))
    base, ms = cpu_reference_run(args, steps, max(1, min(args.warmup, 3)), args.cpu_sample_layers)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": max(1, min(args.warmup, 3)), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, args.gpus),
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "NATriuM itself cannot be built here (deal.II/Trilinos/p4est/Boost/MPI absent): this is the oracle's "
                    "restatement of its loop structure on the host cores"}
    print(json.dumps(line), flush=True)


def workload_config(args, n_gpus):
    nd = args.cells * args.order + 1
    return {"workload": f"TGV3D {args.stencil} BGK semi-Lagrangian p={args.order} {args.cells}^3 cells/GPU "
                        f"({nd}^3 DoFs/GPU) x {n_gpus} GPU slab(s) along z",
            "cells_per_gpu": args.cells ** 3, "fe_order": args.order, "stencil": args.stencil, "collision": "BGK_STANDARD" if args.stencil not in ("D2Q25H", "D3Q45") else "BGK_STANDARD f+g quartic Pr=0.71 Sutherland",
            "cfl": 0.4, "mach": 0.05, "y_stretch": args.stretch, "parallelism": f"slab x{n_gpus} (NCCL ghost exchange)" if n_gpus > 1 else "single GPU",
            "l2_policy": "inputs_exceed_l2 (matrix stream per step >> 126 MB L2)"}


# ------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from natrium_b200 import Context, harness
    from natrium_b200.stencils import Stencil

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local)
    uid = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        box = [Context.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]

    Ma = 0.05
    with_g = args.stencil in ("D2Q25H", "D3Q45")      # compressible two-distribution configurations (not the bench line)
    st = Stencil(args.stencil, 1.0 if with_g else math.sqrt(3) / Ma)
    cells = [args.cells, args.cells, args.cells * world]
    L3 = [2 * math.pi, 2 * math.pi, 2 * math.pi * world]
    verts = None
    if args.stretch > 0.0:      # y-graded mesh (TurbulentChannelFlow3D.h:116-123): one weight-pattern set per cell row
        yy = np.arange(cells[1] + 1) / cells[1]
        verts = [L3[0] * np.arange(cells[0] + 1) / cells[0], L3[1] * (yy - args.stretch * np.sin(2 * math.pi * yy) / (2 * math.pi)),
                 L3[2] * np.arange(cells[2] + 1) / cells[2]]
    pb = harness.CartesianProblem(3, cells, args.order, length=L3, verts=verts)
    dt = pb.timestep(st, 0.4)
    nu = 0.01 if with_g else 2 * math.pi            # Re = 1 as in sl_parallel_benchmark_periodic/benchmark.cpp:58-66
    ctx = Context(local, rank, world, uid)
    ctx.set_stencil(st.getDirections(), st.getWeights(), st.getScaling(), st.getSpeedOfSoundSquare())
    part = harness.SlabPartition(pb, st, dt, rank, world)
    ctx.set_layout(part.n_owned, part.n_ghost, with_g)
    from natrium_b200 import _capi
    fmt_code = {"dict": _capi.FORMAT_DICT, "dict-unstaged": _capi.FORMAT_DICT_UNSTAGED, "ell": _capi.FORMAT_ELL}[args.format]
    ctx.set_matrix_format(fmt_code, args.dedup_tol if args.format != "ell" else 0.0)
    numbering = harness.CellNumbering(part) if args.numbering == "cell" else None
    host = numbering if numbering is not None else part       # what the host sees: points, halo plan
    if args.dof_order == "cell" and numbering is None:
        ctx.set_dof_order(part.cell_blocked_order())
    if args.grid == "on" and args.format == "dict":
        ctx.set_dof_grid(*host.grid_coords(), fe_order=args.order)
    t0 = time.perf_counter()
    nnz = harness.upload_streaming_matrix(ctx, pb, part, st, dt, numbering)
    t_asm = time.perf_counter() - t0
    if world > 1:
        ctx.set_halo(*host.halo_plan())
    n = part.n_owned
    Q, D = st.getQ(), st.getD()
    if with_g:
        ctx.set_collision(nu, dt, equilibrium=_capi.QUARTIC_EQUILIBRIUM, with_g=True, gamma=1.4, prandtl=0.71, sutherland=True)
        rho, u = harness.taylor_green_3d(host.owned_points(), st.getSpeedOfSound(), compressible=True, density_numerator=0.1)
        f0, g0 = harness.quartic_equilibrium_distributions(st, rho, 0.1 * u, np.ones(n), 1.4)
        ctx.upload_populations(1, g0)
    else:
        ctx.set_collision(nu, dt)
        rho, u = harness.taylor_green_3d(host.owned_points(), st.getSpeedOfSound())
        f0 = harness.equilibrium_distributions(st, rho, u)
    ctx.upload_populations(0, f0)
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

    def e2e_step(i):
        ctx.step_host(bufs[i & 1].data_ptr(), bufs[(i + 1) & 1].data_ptr(), rho_ptr, u_ptr, args.e2e_chunks)

    e2e_step(0)
    e2e_step(1)
    barrier()
    ctx.timer_start()
    for i in range(e2e_steps):
        e2e_step(i)
    ms_e2e = ctx.timer_stop()
    barrier()
    assert np.isfinite(hmom.numpy()).all() and np.isfinite(bufs[e2e_steps & 1].numpy()).all()
    if with_g:
        ms_e2e = 0.0                # g stays on the device in nb200_step_host: no end-to-end number for f+g
    if world > 1:
        t = torch.tensor([ms_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    e2e_val = None if with_g else n_global * Q * e2e_steps / (ms_e2e * 1e-3) / 1e6
    h2d = Q * n * 8
    d2h = Q * n * 8 + (1 + D) * n * 8

    # ---- roofline of the dominant kernel (the fused stream+collide kernel: one launch per step per GPU)
    peak, peak_src = measured_peak()
    bpd = algorithmic_bytes_per_dof(nnz / n, D, Q, with_g)
    alg_bytes = bpd * n
    kern_ms = ms / args.steps          # N=1: the step IS the kernel launch; N>1 includes the halo exchange
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    info = ctx.matrix_info()
    cell_rows = args.numbering == "cell" or args.dof_order == "cell"
    traffic = ncu_traffic(f"{args.stencil}_p{args.order}_{args.cells}_{args.format}") if cell_rows else None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "dram_frac_of_peak": (traffic / (kern_ms * 1e-3) / 1e9 / peak) if traffic else None,
                "peak_source": peak_src, "algorithmic_bytes_per_dof": bpd, "algorithmic_bytes_per_launch": alg_bytes,
                "kernel": (f"k_stream_collide_f_staged<{D},{Q},BGK>" if ctx.matrix_format_info().get("staged") else f"k_stream_collide_f<{D},{Q},BGK,{args.format}>"),
                "kernel_ms": kern_ms, "frac_of_nominal_8000": achieved / 8000.0,
                "device_format_bytes": info["device_bytes"], "nnz": nnz,
                "note": "frac is defined on the algorithmic bytes B = 12*nnz + populations (SURVEY 8d); the dictionary format moves ~6x fewer "
                        "DRAM bytes (traffic), so frac > 1; the kernel's actual limiter is the L1/shared-memory data pipe (ncu: "
                        "l1tex__data_pipe_lsu_wavefronts 80 % of peak, profiles/r01_fused_staged_v6_ncu_selected.json)", "matrix_format": ctx.matrix_format_info(), "dof_order": args.dof_order, "host_numbering": args.numbering}

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu_base, _ = cpu_reference_run(args, 0, 1, args.cpu_sample_layers, gpu_parity=True)
            pg = cpu_base.get("parity_vs_gpu")
            if pg is not None:
                pg["tolerance"] = 1e-12
                pg["ok"] = bool(pg["max_rel_err"] <= 1e-12 and pg["row_sum_err"] <= 1e-12)
                if not pg["ok"]:
                    print(f"PARITY GATE FAILED: {pg}", file=sys.stderr, flush=True)
        except Exception as ex:    # the baseline is a reported number, never a reason to lose the bench line
            cpu_base = {"value": None, "unit": UNIT, "cores": None, "kind": "port", "sample": f"failed: {ex}"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
                "clocks": clocks,
                "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                        "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps, "chunks": args.e2e_chunks,
                        "api": "nb200_step_host (pinned host buffers in and out every step)"},
                "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu_base,
                "n_dofs_global": n_global, "matrix_assembly_upload_s": t_asm,
                "conserved": [float(x) for x in cons]}
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
