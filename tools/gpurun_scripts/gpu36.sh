A="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
ncu --set full --clock-control none --import-source on -k regex:dmx_frame32 -s 3 -c 1 -o gpurun_out/prof_frame32_v7 -f $A > gpurun_out/ncu_a.log 2>&1
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --workload hea10"
ncu --set full --clock-control none --import-source on -k regex:dmx_stream3 -s 40 -c 2 -o gpurun_out/prof_stream3_v7 -f $B > gpurun_out/ncu_b.log 2>&1
ls -la gpurun_out/*.ncu-rep
echo done
