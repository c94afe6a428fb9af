set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_n2_h2.json 2> gpurun_out/err_n2.log; tail -n 3 gpurun_out/err_n2.log
timeout 300 $TR --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 1 --impl reference > gpurun_out/bench_n2_ref.json 2> gpurun_out/err_n2r.log
timeout 300 $TR --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --workload h2_qem --no-cpu-baseline > gpurun_out/bench_n2_qem.json 2> gpurun_out/err_n2q.log; tail -n 3 gpurun_out/err_n2q.log
timeout 600 $TR --master-port 29514 bench.py --gpus 2 --steps 2 --warmup 3 --workload hea10 --no-cpu-baseline > gpurun_out/bench_n2_hea.json 2> gpurun_out/err_n2h.log; tail -n 3 gpurun_out/err_n2h.log
for f in h2 qem hea; do python -c "
import json; d=json.load(open('gpurun_out/bench_n2_$f.json')); print('$f n=2 value %.5g'%d['value'], d['ms_per_step'], 'e2e %.5g'%d['e2e']['value'], d['e2e'].get('device_sampling',{}).get('value'))"; done
