set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -2
python bench.py > gpurun_out/bench_h2.json 2> gpurun_out/err_h2.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_h2_ref.json 2> gpurun_out/err_h2r.log
python bench.py --steps 20 --warmup 3 --workload h2_qem > gpurun_out/bench_qem.json 2> gpurun_out/err_qem.log
python bench.py --steps 5 --warmup 3 --workload hea10 > gpurun_out/bench_hea10.json 2> gpurun_out/err_hea.log
python bench.py --steps 3 --warmup 3 --workload lih12 > gpurun_out/bench_lih12.json 2> gpurun_out/err_lih.log
A="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$A > gpurun_out/plain_h2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_h2.csv $A > gpurun_out/ncu_h2.log 2>&1
$A > gpurun_out/plain_h2b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dmx_frame32 -s 3 -c 2 -o gpurun_out/prof_frame32_v5 -f $A > gpurun_out/ncu_h2b.log 2>&1
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --workload hea10"
$B > gpurun_out/plain_hea.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 3900 -c 1400 --csv --log-file gpurun_out/launches_hea10.csv $B > gpurun_out/ncu_hea.log 2>&1
$B > gpurun_out/plain_heab.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dmx_stream3 -s 40 -c 3 -o gpurun_out/prof_stream3_v5 -f $B > gpurun_out/ncu_heab.log 2>&1
echo done
