set -x
python bench.py --steps 2 --warmup 3 --no-also --no-cpu-baseline > gpurun_out/plain2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 400 --csv --log-file gpurun_out/r2_launches_h2_sweep_v2.csv python bench.py --steps 2 --warmup 3 --no-also --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
python bench.py --steps 2 --warmup 3 --no-also --no-cpu-baseline > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:frame32s -s 4 -c 2 -o /tmp/r2_prof_frame32s python bench.py --steps 2 --warmup 3 --no-also --no-cpu-baseline > gpurun_out/ncu3.log 2>&1
ncu -i /tmp/r2_prof_frame32s.ncu-rep --page raw --csv > gpurun_out/r2_ncu_raw_frame32s.csv 2>/dev/null
ncu -i /tmp/r2_prof_frame32s.ncu-rep --page source --csv --launch-skip 1 --launch-count 1 > gpurun_out/r2_ncu_source_frame32s.csv 2>/dev/null
python tools/ncu_raw_summary.py gpurun_out/r2_ncu_raw_frame32s.csv pipe_ shfl
ls -la gpurun_out | tail -8; du -sh gpurun_out
