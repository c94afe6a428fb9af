set -x
timeout 1800 python -m pytest tests -q -m gpu -x --durations=6 2>&1 | tail -14
timeout 900 python bench.py > gpurun_out/r2_bench_v8.json 2> gpurun_out/r2_bench_v8.err; echo bench_rc=$?; tail -3 gpurun_out/r2_bench_v8.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_v8.json'))
print('H2 value %.4g iso %.4g e2e %.4g sync %.4g ms %.5f parity %s cpu %.4g frac %.3f kern %s' % (d['value'], d['isolated']['value'], d['e2e']['value'], d['e2e'].get('synchronous',{}).get('value',0), d['ms_per_step'], d['parity_max_abs_err'], d['cpu_baseline']['value'], d['roofline']['frac'], d['roofline']['kernel'][:22]))
for k,v in d.get('also',{}).items():
    if 'error' in v: print(k, v); continue
    print(k, 'value %.5g ms %.4f e2e %.4g parity %s frac %.3f fexec %s cpu %s wall %s' % (v['value'], v['ms_per_step'], v['e2e']['value'], v['parity_max_abs_err'], v['roofline']['frac'], v['roofline'].get('frac_executed'), (v.get('cpu_baseline') or {}).get('value'), v['wall_s']))
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r2_ref_v8.json 2> gpurun_out/r2_ref_v8.err; echo ref_rc=$?; cut -c1-200 gpurun_out/r2_ref_v8.json
python -c "import __graft_entry__ as g; g.smoke()"
