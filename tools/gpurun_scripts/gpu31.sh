set -x
A="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --workload h2_qem"
$A > gpurun_out/plain_qem.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dmx_frame32 -s 3 -c 1 -o gpurun_out/prof_frame32_qem -f $A > gpurun_out/ncu_qem.log 2>&1
echo done
