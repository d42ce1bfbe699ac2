#!/bin/bash
# round 2, call G (1 GPU): the whole GPU suite (filter tests new), then a quick config-2 timing
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r2g_pytest.log 2>&1
tail -8 gpurun_out/r2g_pytest.log
bash tools/gpu_bench_quick.sh --configs ""
