#!/bin/bash
# round 2, call C: ncu of the grid kernel on config 2
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stream_collide_f_grid -s 2 -c 1 -o gpurun_out/r2c_grid python bench.py --steps 3 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2c_ncu_grid.log 2>&1
tail -2 gpurun_out/r2c_ncu_grid.log | cut -c1-300
