set -x
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
for opts in "--opt stream_chunk=64" "--opt tile_qubits=6 --opt stream_chunk=64" "--opt tile_qubits=6 --opt stream_chunk=16" "" ; do
timeout 300 python bench.py --steps 3 --warmup 3 --workload hea10 --no-cpu-baseline --no-e2e $opts > gpurun_out/bench_hea_x.json 2> gpurun_out/err.log; python -c "
import json; d=json.load(open('gpurun_out/bench_hea_x.json')); print('HEA [$opts] %.5g circuits/s'%d['value'], d['ms_per_step'], d['roofline']['frac'], d['config']['plan'], d['gpu_launches'])"; tail -3 gpurun_out/err.log
done
