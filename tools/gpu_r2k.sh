#!/bin/bash
# round 2, call K (1 GPU): box stores of the grid copy -- grid / host-step / conservation tests, then config-2 timing
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x -k "grid or step_host or conserved or fused_step_matches_oracle_per_step or shim or filtered" 2>&1 | tail -25 ) > gpurun_out/r2k_pytest.log 2>&1
grep -E "passed|failed|FAILED|Error" gpurun_out/r2k_pytest.log | head -20
if grep -q "failed" gpurun_out/r2k_pytest.log; then
  timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "grid_kernels_selected and tgv3d_d3q19_small and lex" > gpurun_out/r2k_sanitizer.log 2>&1
  grep -E "=========|at 0x|by thread|Invalid|illegal|Illegal" gpurun_out/r2k_sanitizer.log | head -40
else
  bash tools/gpu_bench_quick.sh --configs ""
fi
