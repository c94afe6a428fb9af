P=$PWD/vqe_errormitigation_b200
run() { name=$1; shift; env "$@" python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-e2e $EXTRA > gpurun_out/x_$name.json 2>gpurun_out/x_$name.err; python -c "
import json; d=json.load(open('gpurun_out/x_$name.json')); print('$name value %.5g us %.2f'%(d['value'], d['ms_per_step']*1e3))"; }
EXTRA="" run base A=1
EXTRA="" run uncapped DMX_F32_UNCAPPED=1
EXTRA="" run mb4 DMX_LIB=$P/libdmx_mb4.so
EXTRA="" run mb4_uncapped DMX_LIB=$P/libdmx_mb4.so DMX_F32_UNCAPPED=1
EXTRA="" run mb5 DMX_LIB=$P/libdmx_mb5.so
EXTRA="" run mb5_uncapped DMX_LIB=$P/libdmx_mb5.so DMX_F32_UNCAPPED=1
EXTRA="--opt seg_cpb=4" run cpb4 A=1
EXTRA="--opt seg_cpb=4" run cpb4_uncapped DMX_F32_UNCAPPED=1
EXTRA="--opt seg_cpb=4" run cpb4_mb5 DMX_LIB=$P/libdmx_mb5.so
EXTRA="--workload h2_qem" run qem_base A=1
EXTRA="--workload h2_qem" run qem_mb4 DMX_LIB=$P/libdmx_mb4.so
EXTRA="--workload h2_qem" run qem_uncapped DMX_F32_UNCAPPED=1
echo done
