set -x
timeout 1800 python -m pytest tests -q -m gpu -x --durations=6 2>&1 | tail -14
timeout 900 python bench.py > gpurun_out/r2_bench_v7.json 2> gpurun_out/r2_bench_v7.err; echo bench_rc=$?; tail -3 gpurun_out/r2_bench_v7.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_v7.json'))
print('H2 value %.4g iso %.4g e2e %.4g sync %.4g ms %.5f parity %s cpu %.4g frac %.3f kern %s' % (d['value'], d['isolated']['value'], d['e2e']['value'], d['e2e'].get('synchronous',{}).get('value',0), d['ms_per_step'], d['parity_max_abs_err'], d['cpu_baseline']['value'], d['roofline']['frac'], d['roofline']['kernel'][:22]))
for k,v in d.get('also',{}).items():
    if 'error' in v: print(k, v); continue
    print(k, 'value %.5g ms %.4f e2e %.4g parity %s frac %.3f fexec %s cpu %s wall %s' % (v['value'], v['ms_per_step'], v['e2e']['value'], v['parity_max_abs_err'], v['roofline']['frac'], v['roofline'].get('frac_executed'), (v.get('cpu_baseline') or {}).get('value'), v['wall_s']))
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r2_ref_v7.json 2> gpurun_out/r2_ref_v7.err; echo ref_rc=$?; cut -c1-200 gpurun_out/r2_ref_v7.json
python -c "import __graft_entry__ as g; g.smoke()"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dmx_frame16t -s 10 -c 1 -o /tmp/r2_prof_f16ts python bench.py --workload h2_sweep --no-also --no-cpu-baseline --no-e2e --steps 20 --warmup 5 > gpurun_out/ncu_f16ts.log 2>&1
ncu -i /tmp/r2_prof_f16ts.ncu-rep --page raw --csv > gpurun_out/r2_ncu_raw_f16ts.csv 2>/dev/null
ncu -i /tmp/r2_prof_f16ts.ncu-rep --page source --csv > gpurun_out/r2_ncu_source_f16ts.csv 2>/dev/null
python tools/ncu_raw_summary.py gpurun_out/r2_ncu_raw_f16ts.csv
