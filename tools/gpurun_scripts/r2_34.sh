set -x
timeout 900 python -m pytest tests/test_sampler.py tests/test_gpu_dropin.py -q -x -k "sampl or mitig or qp" 2>&1 | tail -3
python tools/qem_breakdown.py 2>&1 | tail -12
python bench.py --workload h2_qem --steps 20 --warmup 5 --no-cpu-baseline --no-also 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('qem value %.4g est %.4g ms_est %.4f devsamp %.4g' % (d['value'], d['estimator']['value'], d['estimator']['ms_per_estimate'], d['e2e']['device_sampling']['value']))"
