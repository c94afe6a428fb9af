timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_sampler.py tests/test_gpu_dropin.py -q -m gpu -x -k "not streaming and not eleven and not ten_qubit and not twelve and not similarity" 2>&1 | tail -3
for w in h2_sweep h2_qem; do python bench.py --steps 200 --warmup 10 --no-cpu-baseline --workload $w > gpurun_out/x_$w.json 2>gpurun_out/x_$w.err; python -c "
import json; d=json.load(open('gpurun_out/x_$w.json')); print('$w value %.5g us %.2f e2e %.5g'%(d['value'], d['ms_per_step']*1e3, d['e2e']['value']), d['e2e'].get('device_sampling',{}).get('value'))"; done
