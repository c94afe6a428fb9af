set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
python bench.py > gpurun_out/bench_h2.json 2> gpurun_out/err_h2.log; tail -c 300 gpurun_out/bench_h2.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_h2_ref.json 2> gpurun_out/err_h2r.log; tail -c 500 gpurun_out/bench_h2_ref.json
python bench.py --steps 20 --warmup 3 --workload h2_qem > gpurun_out/bench_qem.json 2> gpurun_out/err_qem.log; tail -c 300 gpurun_out/bench_qem.json
python bench.py --steps 5 --warmup 3 --workload hea10 > gpurun_out/bench_hea10.json 2> gpurun_out/err_hea.log; tail -c 300 gpurun_out/bench_hea10.json
python bench.py --steps 3 --warmup 3 --workload lih12 > gpurun_out/bench_lih12.json 2> gpurun_out/err_lih.log; tail -c 300 gpurun_out/bench_lih12.json
