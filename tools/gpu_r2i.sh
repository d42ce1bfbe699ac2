#!/bin/bash
# round 2, call I (1 GPU): occupancy / pass-capacity variants of the grid kernel (L1 left for the weight patterns), 24^3 cells
mkdir -p gpurun_out
for v in "" occ4 occ4cap2048 occ5cap1280 occ3cap2048; do
  if [ -n "$v" ]; then export NB200_LIB=$PWD/natrium_b200/variants/lib_$v.so; else unset NB200_LIB; fi
  python bench.py --steps 300 --warmup 5 --no-cpu-baseline --e2e-steps 1 --configs "" --no-gates --cells 24 2> gpurun_out/r2i_$v.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
g=d['roofline']['grid']
print('variant %-14s ms/step %.4f  value %.0f  passes %d cap %d' % ('${v:-default}', d['ms_per_step'], d['value'], g['passes'], g['pass_capacity']))
"
done
