python bench.py > gpurun_out/bench_h2_v9.json 2> gpurun_out/err_h2.log
python bench.py --steps 20 --warmup 3 --workload h2_qem --no-cpu-baseline > gpurun_out/bench_qem_v9.json 2> gpurun_out/err_qem.log
for f in h2 qem; do python -c "
import json; d=json.load(open('gpurun_out/bench_${f}_v9.json')); print('$f value %.5g'%d['value'], d['ms_per_step'], 'e2e %.5g'%d['e2e']['value'], d['roofline']['frac'], d['config']['timing'][:80], d['e2e'].get('transfer','')[:40])"; done
