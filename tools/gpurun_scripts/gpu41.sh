ncu --set full --clock-control none --import-source on -k regex:dmx_qp_sample -s 5 -c 1 -o gpurun_out/prof_sampler_v8 -f python tools/qem_breakdown.py > gpurun_out/ncu_s.log 2>&1
ls -la gpurun_out/prof_sampler_v8.ncu-rep
