set -x
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dmx_frame16t -s 10 -c 1 -o /tmp/r2_prof_f16t python bench.py --workload h2_sweep --no-also --no-cpu-baseline --no-e2e --steps 20 --warmup 5 > gpurun_out/ncu_f16t.log 2>&1
ncu -i /tmp/r2_prof_f16t.ncu-rep --page raw --csv > gpurun_out/r2_ncu_raw_f16t.csv 2>/dev/null
ncu -i /tmp/r2_prof_f16t.ncu-rep --page source --csv > gpurun_out/r2_ncu_source_f16t.csv 2>/dev/null
python tools/ncu_raw_summary.py gpurun_out/r2_ncu_raw_f16t.csv
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_h2_sweep_v3.csv python bench.py --workload h2_sweep --no-also --no-cpu-baseline --no-e2e --steps 20 --warmup 5 > gpurun_out/ncu_f16t_launches.log 2>&1
python tools/summarize_launches.py gpurun_out/r2_launches_h2_sweep_v3.csv 2>&1 | tail -15
