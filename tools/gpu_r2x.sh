#!/bin/bash
# round 2, very last call (1 GPU): k_stream_grid with a register target of 4 CTAs/SM -- configs c3 / c5, then the stream tests
mkdir -p gpurun_out
timeout 170 python bench.py --steps 300 --warmup 5 --no-cpu-baseline --e2e-steps 1 --configs c5,c3 --no-gates > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2x_bench.json').read().strip().splitlines()[-1])
    print('headline %.4f' % d['ms_per_step'])
    for k,v in (d.get('configs') or {}).items(): print(k, v.get('ms_per_step'), v.get('error'))
except Exception as ex: print('parse', ex)
PY
timeout 80 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "(stream_matches_oracle and grid) or thermal_channel or lid_driven" 2>&1 | tail -3
