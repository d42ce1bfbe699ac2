#!/bin/bash
# multi-rank parity worker under a hard timeout, progress kept in gpurun_out/multirank.log; args: world size, case list
mkdir -p gpurun_out
export MULTIRANK_CASES="$2"
export PYTHONUNBUFFERED=1
timeout ${3:-240} python -m torch.distributed.run --nnodes=1 --nproc-per-node=$1 --master-addr 127.0.0.1 --master-port 29531 tests/multirank_check.py > gpurun_out/multirank.log 2>&1
echo "exit $?" >> gpurun_out/multirank.log
grep -v "^W\|^\*\*\*" gpurun_out/multirank.log | tail -25
