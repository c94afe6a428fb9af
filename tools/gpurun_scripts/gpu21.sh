set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_h2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dmx_frame32 -s 3 -c 2 -o gpurun_out/prof_frame32 -f $CMD > gpurun_out/ncu_h2.log 2>&1
tail -n 2 gpurun_out/ncu_h2.log | cut -c1-300
