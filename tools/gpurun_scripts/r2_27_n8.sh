set -x
nproc; python -c "import os; print(len(os.sched_getaffinity(0)))"; lscpu | grep -i -E "socket|numa|model name" | head -8
for g in 0 1 2 3 4 5 6 7; do nvidia-smi -i $g --query-gpu=pci.bus_id --format=csv,noheader; done
cat /sys/devices/system/node/online
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 tools/e2e_scaling.py quick 2>&1 | grep -E "zero_copy|pinning"
DMX_BENCH_NUMA=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 tools/e2e_scaling.py quick 2>&1 | grep -E "zero_copy|pinning"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/r2_bench_n8_v3.json 2> gpurun_out/r2_bench_n8_v3.err; echo rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n8_v3.json'))
print('N=%d H2 value %.4g iso %.4g e2e %.4g sync %.4g ms %.5f' % (d['n_gpus'], d['value'], d['isolated']['value'], d['e2e']['value'], d['e2e'].get('synchronous',{}).get('value',0), d['ms_per_step']), d['config'].get('host_cores_per_rank'), d['config'].get('rank0_numa_node'))
for k,v in d.get('also',{}).items():
    print(k, 'value %.5g ms %.4f e2e %.4g parity %s' % (v['value'], v['ms_per_step'], v['e2e']['value'], v['parity_max_abs_err']))
    if 'estimator' in v: print('  estimator', v['estimator']['value'], v['estimator']['ms_per_estimate'])
PY
