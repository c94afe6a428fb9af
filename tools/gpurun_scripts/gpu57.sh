python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload h2_qem > gpurun_out/bench_qem_v10.json 2> gpurun_out/final_qem.err; tail -3 gpurun_out/final_qem.err; python -c "
import json; d=json.load(open('gpurun_out/bench_qem_v10.json')); print('qem %.4g e2e %.4g'%(d['value'], d['e2e']['value']), d['dedup'])"
