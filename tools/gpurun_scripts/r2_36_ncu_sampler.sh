set -x
python tools/qem_breakdown.py > gpurun_out/plain5.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:qp_sample -s 5 -c 1 -o /tmp/r2_prof_sampler python tools/qem_breakdown.py > gpurun_out/ncu5.log 2>&1
ncu -i /tmp/r2_prof_sampler.ncu-rep --page raw --csv > gpurun_out/r2_ncu_raw_sampler.csv 2>/dev/null
ncu -i /tmp/r2_prof_sampler.ncu-rep --page source --csv > gpurun_out/r2_ncu_source_sampler.csv 2>/dev/null
python tools/ncu_raw_summary.py gpurun_out/r2_ncu_raw_sampler.csv
