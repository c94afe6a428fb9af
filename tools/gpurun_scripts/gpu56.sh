python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/final_h2.json 2> gpurun_out/final_h2.err; python -c "
import json; d=json.load(open('gpurun_out/final_h2.json')); print('h2 %.4g e2e %.4g'%(d['value'], d['e2e']['value']))"
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload h2_qem > gpurun_out/final_qem.json 2> gpurun_out/final_qem.err; python -c "
import json; d=json.load(open('gpurun_out/final_qem.json')); print('qem %.4g e2e %.4g ds %.4g'%(d['value'], d['e2e']['value'], d['e2e']['device_sampling']['value']))"
