set -x
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --workload hea10 > gpurun_out/bench_hea10.json 2> gpurun_out/err_hea.log; tail -c 1200 gpurun_out/bench_hea10.json
python bench.py --steps 3 --warmup 3 --workload lih12 > gpurun_out/bench_lih12.json 2> gpurun_out/err_lih.log; tail -c 1200 gpurun_out/bench_lih12.json
CMD="python bench.py --steps 1 --warmup 3 --workload hea10 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain6.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dmx_stream -s 40 -c 3 -o gpurun_out/prof_stream_v3 -f $CMD > gpurun_out/ncu6.log 2>&1
tail -n 3 gpurun_out/ncu6.log
