set -x
python bench.py --steps 20 --warmup 3 --workload h2_qem --no-cpu-baseline > gpurun_out/bench_qem.json 2> gpurun_out/err_qem.log; python -c "
import json; d=json.load(open('gpurun_out/bench_qem.json')); print('QEM %.5g'%d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['device_sampling'])"
A="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$A > gpurun_out/plain_h2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_h2.csv $A > gpurun_out/ncu_h2.log 2>&1
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --workload hea10"
$B > gpurun_out/plain_hea.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 3900 -c 1400 --csv --log-file gpurun_out/launches_hea10.csv $B > gpurun_out/ncu_hea.log 2>&1
tail -n 2 gpurun_out/ncu_h2.log gpurun_out/ncu_hea.log
