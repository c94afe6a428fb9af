python tools/zc_check.py 2>&1 | tail -5
for zc in 1 0; do python bench.py --steps 200 --warmup 10 --no-cpu-baseline --opt host_zero_copy=$zc > gpurun_out/x_zc$zc.json 2>gpurun_out/x_zc$zc.err; python -c "
import json; d=json.load(open('gpurun_out/x_zc$zc.json')); print('zc=$zc value %.5g e2e %.5g'%(d['value'], d['e2e']['value']))"; done
timeout 600 python -m pytest tests -q -m gpu -x -k "host or dropin or qpu" 2>&1 | tail -3
