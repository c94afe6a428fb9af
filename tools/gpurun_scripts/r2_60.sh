set -x
timeout 600 python -m pytest tests/test_sampler.py tests/test_gpu_parity.py tests/test_gpu_dropin.py -q -m gpu -x -k "fused or estimator or mitigat or variant or sampl" 2>&1 | tail -8
timeout 300 python bench.py --workload h2_qem --no-also --no-cpu-baseline --no-e2e > gpurun_out/x_qem_fused.json 2> gpurun_out/x_qem_fused.err; echo rc=$?; tail -2 gpurun_out/x_qem_fused.err
python - <<PY
import json
d=json.load(open('gpurun_out/x_qem_fused.json'))
print('qem value %.5g ms %.5f parity %s est %.5g ms_est %.5f' % (d['value'], d['ms_per_step'], d['parity_max_abs_err'], d['estimator']['value'], d['estimator']['ms_per_estimate']))
print(d['estimator'])
PY
python -c "import __graft_entry__ as g; g.smoke()"
