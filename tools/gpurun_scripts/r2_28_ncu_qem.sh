set -x
python bench.py --workload h2_qem --steps 2 --warmup 3 --no-also --no-cpu-baseline > gpurun_out/plain4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:frame32_kernel -s 6 -c 1 -o /tmp/r2_prof_frame32var python bench.py --workload h2_qem --steps 2 --warmup 3 --no-also --no-cpu-baseline > gpurun_out/ncu4.log 2>&1
ncu -i /tmp/r2_prof_frame32var.ncu-rep --page raw --csv > gpurun_out/r2_ncu_raw_frame32var.csv 2>/dev/null
ncu -i /tmp/r2_prof_frame32var.ncu-rep --page source --csv > gpurun_out/r2_ncu_source_frame32var.csv 2>/dev/null
python tools/ncu_raw_summary.py gpurun_out/r2_ncu_raw_frame32var.csv
tail -3 gpurun_out/ncu4.log
