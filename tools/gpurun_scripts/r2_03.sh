set -x
python tools/profile_stream.py hea10 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:stream3 -s 36 -c 18 -o gpurun_out/r2_prof_stream3_v1 python tools/profile_stream.py hea10 > gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/plain.log gpurun_out/ncu.log
ls -la gpurun_out/*.ncu-rep
