set -x
for rows in 32768; do
python bench.py --steps 100 --warmup 5 --no-cpu-baseline --opt host_chunk_rows=$rows > gpurun_out/bench_h2_x.json 2> gpurun_out/err_h2.log; python -c "
import json; d=json.load(open('gpurun_out/bench_h2_x.json')); print('H2 [$rows] %.5g'%d['value'], d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'])"; tail -n 2 gpurun_out/err_h2.log
done
for rows in 32768; do
python bench.py --steps 20 --warmup 3 --workload h2_qem --no-cpu-baseline --opt host_chunk_rows=$rows > gpurun_out/bench_qem_x.json 2> gpurun_out/err_qem.log; python -c "
import json; d=json.load(open('gpurun_out/bench_qem_x.json')); print('QEM [$rows] %.5g'%d['value'], d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], 'dev-sampled %.4g'%d['e2e']['device_sampling']['value'])"; tail -n 2 gpurun_out/err_qem.log
done
