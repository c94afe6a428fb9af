set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29521 bench.py --gpus 4 --steps 50 --warmup 5 > gpurun_out/bench_n4_h2.json 2> gpurun_out/err_n4.log; wc -l gpurun_out/bench_n4_h2.json
timeout 300 $TR --master-port 29522 bench.py --gpus 4 --steps 2 --warmup 1 --impl reference > gpurun_out/bench_n4_ref.json 2> gpurun_out/err_n4r.log; wc -l gpurun_out/bench_n4_ref.json
timeout 300 $TR --master-port 29523 bench.py --gpus 4 --steps 10 --warmup 3 --workload h2_qem --no-cpu-baseline > gpurun_out/bench_n4_qem.json 2> gpurun_out/err_n4q.log
for f in h2 qem; do python -c "
import json; d=json.load(open('gpurun_out/bench_n4_$f.json')); print('$f n=4 value %.5g'%d['value'], d['ms_per_step'], 'e2e %.5g'%d['e2e']['value'], d['e2e'].get('device_sampling',{}).get('value'), d['e2e'].get('device_sampling',{}).get('estimate'))"; done
python -c "
import json; d=json.load(open('gpurun_out/bench_n4_ref.json')); print('ref n=4', d['value'], d['n_gpus'])"
