set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_n8_v4.json 2> gpurun_out/r2_bench_n8_v4.err; echo rc=$?
tail -3 gpurun_out/r2_bench_n8_v4.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n8_v4.json'))
print('N=%d H2 value %.4g iso %.4g e2e %.4g sync %.4g ms %.5f' % (d['n_gpus'], d['value'], d['isolated']['value'], d['e2e']['value'], d['e2e'].get('synchronous',{}).get('value',0), d['ms_per_step']))
for k,v in d.get('also',{}).items():
    print(k, 'value %.5g ms %.4f e2e %.4g parity %s' % (v['value'], v['ms_per_step'], v['e2e']['value'], v['parity_max_abs_err']))
    if 'estimator' in v: print('  estimator', v['estimator']['value'], v['estimator']['ms_per_estimate'])
PY
