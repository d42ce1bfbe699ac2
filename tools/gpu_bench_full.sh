#!/bin/bash
# the bench line exactly as the driver runs it (N = 1), plus the reference arm
mkdir -p gpurun_out
T0=$(date +%s)
python bench.py "$@" > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
echo "bench exit $? in $(( $(date +%s) - T0 )) s"
tail -5 gpurun_out/bench_full.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_full.json'))
print('ms/step %.4f value %.0f frac %.3f e2e %.2f ms (%.0f)' % (d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['e2e']['value']))
print('parity', d.get('parity'))
print('cpu', {k:v for k,v in (d.get('cpu_baseline') or {}).items() if k!='sample'})
for k,v in (d.get('configs') or {}).items():
    if 'error' in v: print(k, 'ERROR', v['error']); continue
    print(k, 'ms %.4f value %.0f frac %.3f launches/step %.1f e2e %s parity %s' % (v['ms_per_step'], v['value'], v['roofline']['frac'], v['gpu_launches_per_step'], (v.get('e2e') or {}).get('ms_per_step'), {kk:vv for kk,vv in (v.get('parity') or {}).items() if kk in ('ok','max_rel_err','row_sum_err','error','kernels','mesh')}))
PY
