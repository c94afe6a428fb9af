set -x
timeout 600 python -m pytest tests/test_sampler.py -q -m gpu -x 2>&1 | tail -8
timeout 300 python bench.py --workload h2_qem --no-also --no-cpu-baseline --no-e2e > gpurun_out/x_qem_fused2.json 2> gpurun_out/x_qem_fused2.err; echo rc=$?; tail -2 gpurun_out/x_qem_fused2.err
python - <<PY
import json
d=json.load(open('gpurun_out/x_qem_fused2.json'))
print('qem value %.5g ms %.5f parity %s est %.5g ms_est %.5f' % (d['value'], d['ms_per_step'], d['parity_max_abs_err'], d['estimator']['value'], d['estimator']['ms_per_estimate']))
PY
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_h2_qem_v3.csv python bench.py --workload h2_qem --no-also --no-cpu-baseline --no-e2e --steps 5 --warmup 3 > gpurun_out/ncu_qem_launches3.log 2>&1
python tools/summarize_launches.py gpurun_out/r2_launches_h2_qem_v3.csv 2>&1 | tail -12
