set -x
nproc; free -g | head -2
timeout 1500 python -m pytest tests -q -m gpu -x --durations=8 2>&1 | tail -25
timeout 600 python bench.py > gpurun_out/r2_bench_v1.json 2> gpurun_out/r2_bench_v1.err; echo bench_rc=$?; tail -5 gpurun_out/r2_bench_v1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_v1.json'))
print('H2 value %.4g e2e %.4g sync %.4g ms %.5f parity %s' % (d['value'], d['e2e']['value'], d['e2e'].get('synchronous',{}).get('value',0), d['ms_per_step'], d['parity_max_abs_err']))
for k,v in d.get('also',{}).items():
    if 'error' in v: print(k, v); continue
    print(k, 'value %.5g ms %.4f e2e %.4g parity %s frac %.3f fexec %s wall %s' % (v['value'], v['ms_per_step'], v['e2e']['value'], v['parity_max_abs_err'], v['roofline']['frac'], v['roofline'].get('frac_executed'), v['wall_s']))
    if 'estimator' in v: print('  estimator', v['estimator'])
    if k=='h2_qem': print('  e2e', {kk:(vv if not isinstance(vv,dict) else vv.get('value')) for kk,vv in v['e2e'].items()})
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_ref_v1.json 2> gpurun_out/r2_ref_v1.err; echo ref_rc=$?; cat gpurun_out/r2_ref_v1.json | cut -c1-400
