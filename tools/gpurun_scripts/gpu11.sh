set -x
timeout 600 python -m pytest tests -q -m gpu -x -k "ten_qubit or twelve_qubit" 2>&1 | tail -3
for opts in "--opt tile_qubits=7" "--opt tile_qubits=6" "--opt tile_qubits=7 --opt remap=0"; do
timeout 600 python bench.py --steps 2 --warmup 3 --workload lih12 --no-cpu-baseline --no-e2e $opts > gpurun_out/bench_lih_x.json 2> gpurun_out/err.log; python -c "
import json; d=json.load(open('gpurun_out/bench_lih_x.json')); print('LIH [$opts] %.5g circuits/s'%d['value'], d['ms_per_step'], d['roofline']['frac'], d['config']['plan'], d['gpu_launches'])"; tail -3 gpurun_out/err.log
done
timeout 600 python bench.py --steps 3 --warmup 3 --workload hea10 --no-cpu-baseline > gpurun_out/bench_hea_x.json 2> gpurun_out/err.log; python -c "
import json; d=json.load(open('gpurun_out/bench_hea_x.json')); print('HEA %.5g circuits/s'%d['value'], d['ms_per_step'], d['roofline']['frac'], d['config']['plan'], d['gpu_launches'], d['e2e'])"; tail -3 gpurun_out/err.log
