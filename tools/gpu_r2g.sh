#!/bin/bash
# round 2, call G (1 GPU): the whole GPU suite (filter + shim tests new, fixed-K grid products), then a quick config-2 timing
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > gpurun_out/r2g_pytest.log 2>&1
grep -E "passed|failed|FAILED|Error" gpurun_out/r2g_pytest.log | head -30
bash tools/gpu_bench_quick.sh --configs ""
