timeout 600 python -m pytest tests/test_sampler.py tests/test_gpu_dropin.py -q -m gpu -x -k "sampler or estimator or mitigation" 2>&1 | tail -3
python tools/qem_breakdown.py 2>&1 | tail -6
python bench.py --steps 20 --warmup 3 --workload h2_qem --no-cpu-baseline > gpurun_out/bench_qem_v8.json 2> gpurun_out/err_qem.log
python -c "
import json; d=json.load(open('gpurun_out/bench_qem_v8.json')); print('qem value %.5g'%d['value'], d['ms_per_step'], 'e2e %.5g'%d['e2e']['value'], d['e2e']['device_sampling'])"
