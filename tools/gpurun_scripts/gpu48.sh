set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29531 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/bench_n8_h2_v8.json 2> gpurun_out/err_n8.log; tail -n 2 gpurun_out/err_n8.log
timeout 300 $TR --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 --workload h2_qem --no-cpu-baseline > gpurun_out/bench_n8_qem_v8.json 2> gpurun_out/err_n8q.log; tail -n 2 gpurun_out/err_n8q.log
for f in h2 qem; do python -c "
import json; d=json.load(open('gpurun_out/bench_n8_${f}_v8.json')); print('$f n=8 value %.5g'%d['value'], d['ms_per_step'], 'e2e %.5g'%d['e2e']['value'], d['e2e'].get('device_sampling',{}))"; done
