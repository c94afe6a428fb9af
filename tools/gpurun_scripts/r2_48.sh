set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -q -m gpu -x -k "thread_per_circuit or grouped or batch_sizes or full_size or async_host" 2>&1 | tail -15
for o in "frame16t=1" "frame16t=0" "frame16t=2"; do
  timeout 300 python bench.py --workload h2_sweep --no-also --no-cpu-baseline --opt $o > gpurun_out/x_f16_$o.json 2> gpurun_out/x_f16_$o.err; echo rc=$?
  python - <<PY
import json
d=json.load(open('gpurun_out/x_f16_$o.json'))
print('$o value %.5g ms %.5f parity %s iso %.5g e2e %.5g sync %.5g frac %.3f kern %s' % (d['value'], d['ms_per_step'], d['parity_max_abs_err'], d['isolated']['value'], d['e2e']['value'], d['e2e'].get('synchronous',{}).get('value',0), d['roofline']['frac'], d['roofline']['kernel'][:24]))
PY
done
timeout 300 python bench.py --workload h2_bonds --no-also --no-cpu-baseline > gpurun_out/x_f16_bonds.json 2> gpurun_out/x_f16_bonds.err; echo rc=$?
python - <<PY
import json
d=json.load(open('gpurun_out/x_f16_bonds.json'))
print('bonds value %.5g ms %.5f parity %s iso %.5g e2e %.5g frac %.3f' % (d['value'], d['ms_per_step'], d['parity_max_abs_err'], d['isolated']['value'], d['e2e']['value'], d['roofline']['frac']))
PY
