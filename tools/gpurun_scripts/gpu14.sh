set -x
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 --workload h2_qem --no-cpu-baseline > gpurun_out/bench_qem.json 2> gpurun_out/err_qem.log; python -c "
import json; d=json.load(open('gpurun_out/bench_qem.json')); print('QEM %.5g'%d['value'], d['ms_per_step'], d['e2e'])"; tail -3 gpurun_out/err_qem.log
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_h2.json 2> gpurun_out/err_h2.log; python -c "
import json; d=json.load(open('gpurun_out/bench_h2.json')); print('H2 %.5g'%d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline'])"
