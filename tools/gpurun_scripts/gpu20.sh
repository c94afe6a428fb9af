set -x
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
for opts in "" "--opt segment_cpb=4" "--opt frame32=0"; do
python bench.py --steps 50 --warmup 5 --no-cpu-baseline $opts > gpurun_out/bench_h2_x.json 2> gpurun_out/err_h2.log; python -c "
import json; d=json.load(open('gpurun_out/bench_h2_x.json')); print('H2 [$opts] %.5g'%d['value'], d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], d['roofline']['frac'])"; tail -n 2 gpurun_out/err_h2.log
done
python bench.py --steps 20 --warmup 3 --workload h2_qem --no-cpu-baseline > gpurun_out/bench_qem_x.json 2> gpurun_out/err_qem.log; python -c "
import json; d=json.load(open('gpurun_out/bench_qem_x.json')); print('QEM %.5g'%d['value'], d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], 'dev-sampled %.4g'%d['e2e']['device_sampling']['value'])"; tail -n 2 gpurun_out/err_qem.log
