#!/bin/bash
# round 2, call A: compute-sanitizer over a reduced GPU suite + ncu of uniform vs y-graded config 2
mkdir -p gpurun_out
K1='fused_step_matches_oracle_per_step or stream_with_wall_hits or thermal_walls or step_host or stream_matches_oracle or lid_driven'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$K1" > gpurun_out/r2a_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/r2a_memcheck.log
K2='(fused_step_matches_oracle_per_step and (d3q19 or d2q25 or d2q9)) or thermal_walls or (step_host and d3q19)'
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$K2" > gpurun_out/r2a_racecheck.log 2>&1
echo "racecheck exit $?" >> gpurun_out/r2a_racecheck.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stream_collide_f_staged -s 2 -c 1 -o gpurun_out/r2a_uniform python bench.py --steps 3 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2a_ncu_uniform.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stream_collide_f_staged -s 2 -c 1 -o gpurun_out/r2a_stretch python bench.py --steps 3 --warmup 1 --no-cpu-baseline --e2e-steps 1 --stretch 0.8 > gpurun_out/r2a_ncu_stretch.log 2>&1
tail -3 gpurun_out/r2a_memcheck.log gpurun_out/r2a_racecheck.log
