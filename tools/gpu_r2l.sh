#!/bin/bash
# round 2, call L (4 GPUs): multi-rank parity on 2x2x1 blocks / 4 slabs, then the 4-GPU bench line
mkdir -p gpurun_out
bash tools/gpu_multirank.sh 4 "staged-d3q19,grid-d3q19-p4,grid-d3q45,block-d3q19,block-d2q25,hoststep-d3q19-p4,grid-walled" 300
cp gpurun_out/multirank.log gpurun_out/multirank4.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node=4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 300 --warmup 5 --e2e-steps 4 > gpurun_out/r2l_bench_n4.json 2> gpurun_out/r2l_bench_n4.err
echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2l_bench_n4.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','n_gpus','e2e','parity_multirank','partition_block'): print(k, d.get(k))
PY
grep -v "^W\|^\*\*\*" gpurun_out/r2l_bench_n4.err | tail -5
