set -x
timeout 900 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_parity.py -q -x -k "grouped or collection or plan_options or evaluate_variants or async or h2" --durations=5 2>&1 | tail -12
timeout 600 python bench.py > gpurun_out/r2_bench_v2.json 2> gpurun_out/r2_bench_v2.err; echo bench_rc=$?; tail -5 gpurun_out/r2_bench_v2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_v2.json'))
print('H2 value %.4g e2e %.4g sync %.4g ms %.5f parity %s' % (d['value'], d['e2e']['value'], d['e2e'].get('synchronous',{}).get('value',0), d['ms_per_step'], d['parity_max_abs_err']))
for k,v in d.get('also',{}).items():
    if 'error' in v: print(k, v); continue
    print(k, 'value %.5g ms %.4f e2e %.4g parity %s frac %.3f fexec %s wall %s' % (v['value'], v['ms_per_step'], v['e2e']['value'], v['parity_max_abs_err'], v['roofline']['frac'], v['roofline'].get('frac_executed'), v['wall_s']))
    if 'estimator' in v: print('  estimator', v['estimator']['value'], v['estimator']['ms_per_estimate'])
    if k=='h2_qem': print('  e2e', {kk:(vv if not isinstance(vv,dict) else vv.get('value')) for kk,vv in v['e2e'].items() if kk in ('value','raw_table','device_sampling','unique_rows')})
PY
