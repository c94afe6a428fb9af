set -x
CMD="python bench.py --steps 1 --warmup 3 --workload hea10 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain6.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dmx_stream3 -s 40 -c 3 -o gpurun_out/prof_stream3_v4 -f $CMD > gpurun_out/ncu6.log 2>&1
tail -n 3 gpurun_out/ncu6.log
