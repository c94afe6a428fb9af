set -x
timeout 900 python -m pytest tests/test_sampler.py tests/test_gpu_dropin.py tests/test_gpu_parity.py -q -x -k "sampl or mitig or qp or variant or evaluate" 2>&1 | tail -5
python tools/qem_breakdown.py 2>&1 | tail -6
python bench.py --workload h2_qem --steps 20 --warmup 5 --no-cpu-baseline --no-also 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('qem value %.4g est %.4g ms_est %.4f devsamp %.4g' % (d['value'], d['estimator']['value'], d['estimator']['ms_per_estimate'], d['e2e']['device_sampling']['value']))"
python bench.py --workload swap7_qem --steps 5 --warmup 3 --no-cpu-baseline --no-also 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('swap7 value %.4g est %.4g parity %s' % (d['value'], d['estimator']['value'], d['parity_max_abs_err']))"
