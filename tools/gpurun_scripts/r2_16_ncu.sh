set -x
python tools/profile_stream.py hea10 > gpurun_out/plain1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:stream3 -s 36 -c 18 -o /tmp/r2_prof_stream3_v2 python tools/profile_stream.py hea10 > gpurun_out/ncu1.log 2>&1
ncu -i /tmp/r2_prof_stream3_v2.ncu-rep --page raw --csv > gpurun_out/r2_ncu_raw_stream3_v2.csv 2>/dev/null
ncu -i /tmp/r2_prof_stream3_v2.ncu-rep --page source --csv --launch-skip 2 --launch-count 1 > gpurun_out/r2_ncu_source_stream3_pass2_bulk.csv 2>/dev/null
python bench.py --steps 2 --warmup 3 --no-also --no-cpu-baseline > gpurun_out/plain2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 400 --csv --log-file gpurun_out/r2_launches_h2_sweep.csv python bench.py --steps 2 --warmup 3 --no-also --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
python bench.py --steps 2 --warmup 3 --no-also --no-cpu-baseline > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:frame32 -s 4 -c 2 -o gpurun_out/r2_prof_frame32_v2 python bench.py --steps 2 --warmup 3 --no-also --no-cpu-baseline > gpurun_out/ncu3.log 2>&1
python bench.py --workload h2_qem --steps 2 --warmup 3 --no-also --no-cpu-baseline > gpurun_out/plain4.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 300 --csv --log-file gpurun_out/r2_launches_h2_qem.csv python bench.py --workload h2_qem --steps 2 --warmup 3 --no-also --no-cpu-baseline > gpurun_out/ncu4.log 2>&1
ls -la gpurun_out/ | tail -12; du -sh gpurun_out
