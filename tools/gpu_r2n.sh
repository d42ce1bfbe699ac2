#!/bin/bash
# round 2, call N (4 GPUs): device timeline of a multi-rank step (NB200_TRACE=1), pass capacity 1536
mkdir -p gpurun_out
export NB200_TRACE=1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 300 --warmup 5 --e2e-steps 1 --no-gates --no-block-partition --no-cpu-baseline 2> gpurun_out/r2n.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('N=4 cap1536: ms/step %.4f value %.0f launches %d' % (d['ms_per_step'], d['value'], d['gpu_launches']))
"
grep "nb200 trace" gpurun_out/r2n.err
