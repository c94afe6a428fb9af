set -x
timeout 300 python tools/trajectory_rate.py > gpurun_out/r2_trajectory_rate.jsonl 2> gpurun_out/r2_trajectory_rate.err; echo rc=$?; cat gpurun_out/r2_trajectory_rate.jsonl; tail -3 gpurun_out/r2_trajectory_rate.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dmx_tile_kernel -s 4 -c 1 -o /tmp/r2_prof_tile7 python bench.py --workload swap7_qem --no-also --no-cpu-baseline --no-e2e --steps 3 --warmup 3 > gpurun_out/ncu_tile7.log 2>&1
ncu -i /tmp/r2_prof_tile7.ncu-rep --page raw --csv > gpurun_out/r2_ncu_raw_tile7.csv 2>/dev/null
ncu -i /tmp/r2_prof_tile7.ncu-rep --page source --csv > gpurun_out/r2_ncu_source_tile7.csv 2>/dev/null
python tools/ncu_raw_summary.py gpurun_out/r2_ncu_raw_tile7.csv
