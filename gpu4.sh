set -x
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -8
timeout 300 python bench.py --steps 3 --warmup 3 --workload hea10 --no-cpu-baseline > gpurun_out/bench_hea.json 2> gpurun_out/err.log; tail -c 1500 gpurun_out/bench_hea.json; tail -3 gpurun_out/err.log
timeout 300 python bench.py --steps 3 --warmup 3 --workload hea10 --no-cpu-baseline --no-e2e > gpurun_out/plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:dmx_tile -s 30 -c 2 -o gpurun_out/prof_tile python bench.py --steps 3 --warmup 3 --workload hea10 --no-cpu-baseline --no-e2e > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
