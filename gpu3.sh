set -x
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -8
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_h2_cpb8.json 2> gpurun_out/err.log; tail -c 1200 gpurun_out/bench_h2_cpb8.json; tail -3 gpurun_out/err.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --opt segment_cpb=4 > gpurun_out/bench_h2_cpb4.json 2> gpurun_out/err.log; tail -c 1200 gpurun_out/bench_h2_cpb4.json; tail -3 gpurun_out/err.log
timeout 300 python bench.py --steps 20 --warmup 3 --workload h2_qem --no-cpu-baseline > gpurun_out/bench_qem.json 2> gpurun_out/err.log; tail -c 1500 gpurun_out/bench_qem.json; tail -3 gpurun_out/err.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:dmx_frame -s 3 -c 2 -o gpurun_out/prof_frame python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
