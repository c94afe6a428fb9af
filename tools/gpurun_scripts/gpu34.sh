set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_sampler.py -q -m gpu -x -k "not streaming and not eleven and not ten_qubit and not twelve" 2>&1 | tail -3
python bench.py --no-cpu-baseline > gpurun_out/bench_h2_v7.json 2> gpurun_out/err_h2.log
python bench.py --steps 20 --warmup 3 --workload h2_qem --no-cpu-baseline > gpurun_out/bench_qem_v7.json 2> gpurun_out/err_qem.log
for f in h2 qem; do python -c "
import json; d=json.load(open('gpurun_out/bench_${f}_v7.json')); print('$f value %.5g'%d['value'], d['ms_per_step'], 'e2e %.5g'%d['e2e']['value'], d['roofline']['frac'])"; done
A="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_h2_v7.csv $A > gpurun_out/ncu_h2.log 2>&1
echo done
