#!/bin/bash
# round 2, final call (1 GPU): whole GPU suite, ncu capture of the final headline kernel, the bench line as the driver runs it,
# the reference arm, and the worst-case (one pattern per row) side line
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -12 ) > gpurun_out/r2z_pytest.log 2>&1
grep -E "passed|failed|FAILED|Error" gpurun_out/r2z_pytest.log | head -20
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stream_collide_f_grid -s 2 -c 1 -o gpurun_out/r2z_grid python bench.py --steps 3 --warmup 1 --no-cpu-baseline --e2e-steps 1 --configs "" --no-gates > gpurun_out/r2z_ncu_grid.log 2>&1
# the bench line quotes the capture of THIS build: summarise it into profiles/ (copied back through gpurun_out/)
python tools/ncu_summary.py gpurun_out/r2z_grid.ncu-rep gpurun_out/r02_grid_v4_ncu_selected.json > gpurun_out/r2z_ncu_summary.txt 2>&1 && cp gpurun_out/r02_grid_v4_ncu_selected.json profiles/ && \
  python tools/update_traffic.py D3Q19_p4_32_grid profiles/r02_grid_v4_ncu_selected.json "profiles/r02_grid_v4_ncu_selected.json (final gpurun call of round 2: k_stream_collide_f_grid<3,19,BGK> with box stores, ncu --set full --clock-control none, taken right before this bench run on the same box)"
T0=$(date +%s)
python bench.py > gpurun_out/r2z_bench_final.json 2> gpurun_out/r2z_bench_final.err
echo "bench exit $? in $(( $(date +%s) - T0 )) s"
T0=$(date +%s)
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2z_bench_reference.json 2> gpurun_out/r2z_bench_reference.err
echo "reference arm exit $? in $(( $(date +%s) - T0 )) s"
timeout 600 python bench.py --steps 100 --warmup 3 --no-cpu-baseline --e2e-steps 1 --configs "" --cells 24 --row-noise 1e-9 --dedup-tol 0 --no-gates > gpurun_out/r2z_side_rownoise.json 2> gpurun_out/r2z_side_rownoise.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2z_bench_final.json').read().strip().splitlines()[-1])
print('ms/step %.4f value %.0f frac %.3f frac_dram %s e2e %.2f ms (%.0f)' % (d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline'].get('frac_on_dram_bytes'), d['e2e']['ms_per_step'], d['e2e']['value']))
print('parity', d.get('parity'))
print('cpu', {k:v for k,v in (d.get('cpu_baseline') or {}).items() if k!='sample'})
for k,v in (d.get('configs') or {}).items():
    if 'error' in v: print(k, 'ERROR', v['error']); continue
    print(k, 'ms %.4f value %.0f frac %.3f launches/step %.1f e2e %s parity %s' % (v['ms_per_step'], v['value'], v['roofline']['frac'], v['gpu_launches_per_step'], (v.get('e2e') or {}).get('ms_per_step'), {kk:vv for kk,vv in (v.get('parity') or {}).items() if kk in ('ok','max_rel_err','error')}))
try:
    r=json.loads(open('gpurun_out/r2z_bench_reference.json').read().strip().splitlines()[-1]); print('reference arm', r.get('value'), r.get('cpu_baseline',{}).get('kind'), r.get('cpu_baseline',{}).get('cores'))
except Exception as ex: print('reference arm parse', ex)
try:
    w=json.loads(open('gpurun_out/r2z_side_rownoise.json').read().strip().splitlines()[-1]); print('row noise: ms %.4f value %.0f frac %.3f patterns %s rows %d' % (w['ms_per_step'], w['value'], w['roofline']['frac'], w['roofline']['matrix_format']['patterns'], w['roofline']['grid']['box_rows']))
except Exception as ex: print('row noise parse', ex)
PY
tail -3 gpurun_out/r2z_bench_final.err gpurun_out/r2z_side_rownoise.err
