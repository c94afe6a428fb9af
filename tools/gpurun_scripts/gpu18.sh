set -x
timeout 900 python -m pytest tests -q -m gpu -x -k "streaming or ten_qubit or twelve_qubit" 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 --workload hea10 --no-cpu-baseline --no-e2e > gpurun_out/bench_hea_x.json 2> gpurun_out/err_hea.log; python -c "
import json; d=json.load(open('gpurun_out/bench_hea_x.json')); print('HEA %.5g circuits/s'%d['value'], d['ms_per_step'], d['roofline']['frac'])"; tail -n 2 gpurun_out/err_hea.log
python bench.py --steps 2 --warmup 3 --workload lih12 --no-cpu-baseline --no-e2e > gpurun_out/bench_lih_x.json 2> gpurun_out/err_lih.log; python -c "
import json; d=json.load(open('gpurun_out/bench_lih_x.json')); print('LIH %.5g circuits/s'%d['value'], d['ms_per_step'], d['roofline']['frac'])"; tail -n 2 gpurun_out/err_lih.log
