set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -3
python bench.py > gpurun_out/bench_h2_v8.json 2> gpurun_out/err_h2.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_h2_ref_v8.json 2> gpurun_out/err_h2r.log
python bench.py --steps 20 --warmup 3 --workload h2_qem > gpurun_out/bench_qem_v8.json 2> gpurun_out/err_qem.log
python bench.py --steps 5 --warmup 3 --workload hea10 > gpurun_out/bench_hea10_v8.json 2> gpurun_out/err_hea.log
python bench.py --steps 3 --warmup 3 --workload lih12 > gpurun_out/bench_lih12_v8.json 2> gpurun_out/err_lih.log
A="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_h2_v8.csv $A > gpurun_out/ncu_h2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_qem_v8.csv $A --workload h2_qem > gpurun_out/ncu_qem.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dmx_frame32 -s 3 -c 1 -o gpurun_out/prof_frame32_v8 -f $A --no-e2e > gpurun_out/ncu_h2b.log 2>&1
for f in h2 qem hea10 lih12; do python -c "
import json; d=json.load(open('gpurun_out/bench_${f}_v8.json')); print('$f value %.5g'%d['value'], d['ms_per_step'], 'e2e %.5g'%d['e2e']['value'], d['roofline']['frac'], d.get('cpu_baseline',{}).get('value'))"; done
echo done
