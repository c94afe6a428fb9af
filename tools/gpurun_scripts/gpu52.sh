set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -3
python tools/host_call_scaling.py 2>&1 | tail -6
python bench.py > gpurun_out/bench_h2_v10.json 2> gpurun_out/err_h2.log
python -c "
import json; d=json.load(open('gpurun_out/bench_h2_v10.json')); print('h2 value %.5g'%d['value'], d['ms_per_step'], 'e2e %.5g'%d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline'])"
