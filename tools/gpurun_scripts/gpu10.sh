set -x
CMD7="python bench.py --steps 1 --warmup 3 --workload hea10 --no-cpu-baseline --no-e2e --opt stream_chunk=64"
CMD6="python bench.py --steps 1 --warmup 3 --workload hea10 --no-cpu-baseline --no-e2e --opt stream_chunk=64 --opt tile_qubits=6"
$CMD7 > gpurun_out/plain7.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dmx_stream -s 30 -c 2 -o gpurun_out/prof_stream7 -f $CMD7 > gpurun_out/ncu7.log 2>&1
$CMD6 > gpurun_out/plain6.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dmx_stream -s 60 -c 2 -o gpurun_out/prof_stream6 -f $CMD6 > gpurun_out/ncu6.log 2>&1
tail -3 gpurun_out/ncu7.log gpurun_out/ncu6.log
