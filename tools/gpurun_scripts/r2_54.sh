set -x
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches_h2_qem_v2.csv python bench.py --workload h2_qem --no-also --no-cpu-baseline --no-e2e --steps 20 --warmup 3 > gpurun_out/ncu_qem_launches.log 2>&1
python tools/summarize_launches.py gpurun_out/r2_launches_h2_qem_v2.csv 2>&1 | tail -14
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dmx_frame16v -s 6 -c 1 -o /tmp/r2_prof_f16v python bench.py --workload h2_qem --no-also --no-cpu-baseline --no-e2e --steps 20 --warmup 3 > gpurun_out/ncu_f16v.log 2>&1
ncu -i /tmp/r2_prof_f16v.ncu-rep --page raw --csv > gpurun_out/r2_ncu_raw_f16v.csv 2>/dev/null
ncu -i /tmp/r2_prof_f16v.ncu-rep --page source --csv > gpurun_out/r2_ncu_source_f16v.csv 2>/dev/null
python tools/ncu_raw_summary.py gpurun_out/r2_ncu_raw_f16v.csv
