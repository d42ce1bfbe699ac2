#!/bin/bash
# round 2, last call (1 GPU): smoke(), and config 5 over 300 and over 2000 steps with the clocks sampled (5.8 ms/step over 2000
# steps against 3.8 over 300: throttling or the run itself?)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
nvidia-smi --query-gpu=timestamp,clocks.sm,power.draw,clocks_throttle_reasons.active --format=csv -lms 250 > gpurun_out/r2y_smi.csv 2>&1 &
SMI=$!
python - <<'PY' 2>&1 | tail -12
import sys, json, time, types
sys.argv = ['bench.py', '--no-gates']
import bench
args = bench.parse()
args.no_gates = True
for steps in (300, 2000, 300):
    args.steps = steps
    import builtins
    # run_side_config caps the steps at 300: lift the cap for this experiment
    src = bench.run_side_config
    c = bench.case_spec('c5', args)
    B = bench.build_product(c, c['cells'], c['length'], grid=args.grid, fmt=args.format, tol=args.dedup_tol, numbering=args.numbering, dof_order=args.dof_order)
    ctx = B['ctx']
    import torch
    def barrier():
        ctx.synchronize(); torch.cuda.synchronize()
    ctx.collide()
    t0 = time.time()
    ms, launches = bench.time_steps(ctx, steps, 3, barrier)
    cons = ctx.conserved()
    rho, u, T, s = ctx.download_moments(want_T=True)
    print(f"c5 steps {steps}: {ms/steps:.4f} ms/step, wall {time.time()-t0:.1f} s, conserved {cons}, T range [{T.min():.4g}, {T.max():.4g}], rho range [{rho.min():.4g}, {rho.max():.4g}]", flush=True)
    ctx.close()
PY
kill $SMI
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2y_smi.csv')) if len(r)>=4][1:]
clk=[int(r[1].split()[0]) for r in rows if r[1].strip().split()[0].isdigit()]
pw=[float(r[2].split()[0]) for r in rows if r[2].strip().split()[0].replace('.','').isdigit()]
print('smi samples', len(rows), 'sm clock min/median/max', min(clk), sorted(clk)[len(clk)//2], max(clk), 'power max', max(pw), 'reasons', sorted(set(r[3].strip() for r in rows)))
PY
