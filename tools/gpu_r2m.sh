#!/bin/bash
# round 2, call M (2 GPUs): does the exchange overlap once the fused kernel leaves shared memory free?  pass capacity 1536 vs 1024 vs serial exchange
mkdir -p gpurun_out
run() {
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 400 --warmup 5 --e2e-steps 1 --no-gates --no-block-partition --no-cpu-baseline 2> gpurun_out/r2m_$1.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1: ms/step %.4f value %.0f passes %d cap %d launches %d' % (d['ms_per_step'], d['value'], d['roofline']['grid']['passes'], d['roofline']['grid']['pass_capacity'], d['gpu_launches']))
"
}
NB200_GRID_CAP=1536 run cap1536
run cap1024_default
NB200_OVERLAP=0 NB200_GRID_CAP=1536 run serial_cap1536
