set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -3
python bench.py > gpurun_out/bench_h2_v6.json 2> gpurun_out/err_h2.log
python bench.py --steps 20 --warmup 3 --workload h2_qem > gpurun_out/bench_qem_v6.json 2> gpurun_out/err_qem.log
python bench.py --steps 5 --warmup 3 --workload hea10 --no-cpu-baseline > gpurun_out/bench_hea10_v6.json 2> gpurun_out/err_hea.log
for f in h2 qem hea10; do python -c "
import json; d=json.load(open('gpurun_out/bench_${f}_v6.json')); print('$f value %.5g'%d['value'], d['ms_per_step'], 'e2e %.5g'%d['e2e']['value'], d['roofline']['frac'])"; done
echo done
