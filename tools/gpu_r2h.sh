#!/bin/bash
# round 2, call H (1 GPU): ncu of the fixed-K grid kernel + launch list, worst-case matrices, sanitizer over the new kernels
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stream_collide_f_grid -s 2 -c 1 -o gpurun_out/r2h_grid python bench.py --steps 3 --warmup 1 --no-cpu-baseline --e2e-steps 1 --configs "" --no-gates > gpurun_out/r2h_ncu_grid.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 2 --configs "" --no-gates > gpurun_out/r2h_launches.log 2>&1
# worst-case matrices: jittered vertices (one weight pattern per row and direction), y-graded mesh, bit-exact dictionary
for v in "--jitter 0.2" "--stretch 0.8" "--dedup-tol 0"; do
  python bench.py --steps 200 --warmup 5 --no-cpu-baseline --e2e-steps 1 --configs "" $v > "gpurun_out/r2h_side_$(echo $v | tr -d ' -.').json" 2>> gpurun_out/r2h_side.err
done
K1='exponential_filter or filtered_step or (grid_kernels_selected and d3q19) or (fused_step_matches_oracle_per_step and grid and (d3q19 or d2q25)) or grid_step_host'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$K1" > gpurun_out/r2h_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/r2h_memcheck.log
K2='(exponential_filter and (d3q19_p2 or d2q25)) or (fused_step_matches_oracle_per_step and grid and (d3q19 or d2q25))'
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$K2" > gpurun_out/r2h_racecheck.log 2>&1
echo "racecheck exit $?" >> gpurun_out/r2h_racecheck.log
tail -3 gpurun_out/r2h_memcheck.log gpurun_out/r2h_racecheck.log
for f in gpurun_out/r2h_side_*.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r=d['roofline']
print(sys.argv[1], 'ms %.4f value %.0f frac %.3f patterns %s grid %s parity %s' % (d['ms_per_step'], d['value'], r['frac'], r['matrix_format']['patterns'], {k:r['grid'][k] for k in ('in_use','box_rows','generic_rows')}, (d.get('parity') or {}).get('max_rel_err')))
PY
done
