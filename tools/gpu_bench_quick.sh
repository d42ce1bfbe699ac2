#!/bin/bash
# quick device-resident timing of config 2 (no CPU baseline, no e2e to speak of); args: extra bench.py flags
mkdir -p gpurun_out
python bench.py --steps 300 --warmup 5 --no-cpu-baseline --e2e-steps 1 "$@" 2> gpurun_out/quick.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms/step %.4f  value %.0f  frac %.3f  clocks %s' % (d['ms_per_step'], d['value'], d['roofline']['frac'], d['clocks']))
"
tail -2 gpurun_out/quick.err
