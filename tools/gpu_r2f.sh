#!/bin/bash
# round 2, call F (2 GPUs): multi-rank parity worker (all cases), then a short 2-GPU bench line (parity_multirank + e2e at N=2)
mkdir -p gpurun_out
bash tools/gpu_multirank.sh 2 "" 400
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 200 --warmup 5 --e2e-steps 5 > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err
echo "bench exit $?"
tail -c 1500 gpurun_out/r2f_bench_n2.json
grep -v "^W\|^\*\*\*" gpurun_out/r2f_bench_n2.err | tail -5
