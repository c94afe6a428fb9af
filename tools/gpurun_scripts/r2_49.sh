set -x
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -q -m gpu -x -k "thread_per_circuit or grouped or batch_sizes or full_size or async_host" 2>&1 | tail -5
for o in "frame16t_threads=64" "frame16t_threads=128" "frame16t_threads=256" "frame16t_threads=64 --opt frame16t=2"; do
  tag=$(echo $o | tr ' =' '__' | tr -d '-')
  timeout 300 python bench.py --workload h2_sweep --no-also --no-cpu-baseline --opt $o > gpurun_out/x_$tag.json 2> gpurun_out/x_$tag.err; echo rc=$?
  python - <<PY
import json
d=json.load(open('gpurun_out/x_$tag.json'))
print('$o value %.5g ms %.5f parity %s iso %.5g e2e %.5g sync %.5g frac %.3f' % (d['value'], d['ms_per_step'], d['parity_max_abs_err'], d['isolated']['value'], d['e2e']['value'], d['e2e'].get('synchronous',{}).get('value',0), d['roofline']['frac']))
PY
done
