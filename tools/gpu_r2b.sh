#!/bin/bash
# round 2, call B: first run of the grid (TMA box) kernels: parity tests, then bench grid on / off
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "grid" > gpurun_out/r2b_pytest_grid.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2b_pytest_grid.log
tail -15 gpurun_out/r2b_pytest_grid.log
timeout 600 python bench.py --steps 300 --warmup 5 --no-cpu-baseline --e2e-steps 3 --grid on > gpurun_out/r2b_bench_grid.json 2> gpurun_out/r2b_bench_grid.err
tail -3 gpurun_out/r2b_bench_grid.err
python -c "
import json
d=json.load(open('gpurun_out/r2b_bench_grid.json'))
print('GRID ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['ms_per_step'])
"
