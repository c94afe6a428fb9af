set -x
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -q -m gpu -x -k "thread_per_circuit or grouped or batch_sizes or full_size or async_host" 2>&1 | tail -5
for o in "frame16t_scaled=1" "frame16t_scaled=0"; do
  timeout 300 python bench.py --workload h2_sweep --no-also --no-cpu-baseline --no-e2e --opt $o > gpurun_out/x_$o.json 2> gpurun_out/x_$o.err; echo rc=$?
  python - <<PY
import json
d=json.load(open('gpurun_out/x_$o.json'))
print('$o value %.5g ms %.5f parity %s iso %.5g frac %.3f' % (d['value'], d['ms_per_step'], d['parity_max_abs_err'], d['isolated']['value'], d['roofline']['frac']))
PY
done
timeout 300 python bench.py --workload h2_sweep --no-also --no-cpu-baseline --no-e2e --steps 500 > gpurun_out/x_k500.json 2> gpurun_out/x_k500.err; python -c "
import json; d=json.load(open('gpurun_out/x_k500.json')); print('K=500 value %.5g ms %.5f' % (d['value'], d['ms_per_step']))"
