set -x
A="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --workload h2_qem"
$A > gpurun_out/plain_qem.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_qem.csv $A > gpurun_out/ncu_qem.log 2>&1
echo done
