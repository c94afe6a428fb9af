set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -q -m gpu -x -k "variant or qp or full_size or mitigat or evaluate" 2>&1 | tail -8
for o in "frame16t=1" "frame16t=0"; do
  timeout 300 python bench.py --workload h2_qem --no-also --no-cpu-baseline --opt $o > gpurun_out/x_qem_$o.json 2> gpurun_out/x_qem_$o.err; echo rc=$?
  python - <<PY
import json
d=json.load(open('gpurun_out/x_qem_$o.json'))
print('$o value %.5g ms %.5f parity %s iso %.5g e2e %.5g est %.5g dedup %.5g frac %.3f' % (d['value'], d['ms_per_step'], d['parity_max_abs_err'], d['isolated']['value'], d['e2e']['value'], d['estimator']['value'], d['dedup']['value'], d['roofline']['frac']))
PY
done
