set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_n2_h2_v8.json 2> gpurun_out/err_n2.log; tail -n 2 gpurun_out/err_n2.log
timeout 300 $TR --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 1 --impl reference > gpurun_out/bench_n2_ref_v8.json 2> gpurun_out/err_n2r.log
timeout 300 $TR --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --workload h2_qem --no-cpu-baseline > gpurun_out/bench_n2_qem_v8.json 2> gpurun_out/err_n2q.log; tail -n 2 gpurun_out/err_n2q.log
for f in h2 qem; do python -c "
import json; d=json.load(open('gpurun_out/bench_n2_${f}_v8.json')); print('$f n=2 value %.5g'%d['value'], d['ms_per_step'], 'e2e %.5g'%d['e2e']['value'], d['e2e'].get('device_sampling',{}).get('value'))"; done
python -c "
import json; d=json.load(open('gpurun_out/bench_n2_ref_v8.json')); print('ref n=2', d['value'], d['n_gpus'])"
