P=$PWD/vqe_errormitigation_b200
run() { name=$1; shift; env "$@" python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --workload hea10 $EXTRA > gpurun_out/x_$name.json 2>gpurun_out/x_$name.err; python -c "
import json; d=json.load(open('gpurun_out/x_$name.json')); print('$name value %.5g ms %.2f frac %.3f'%(d['value'], d['ms_per_step'], d['roofline']['frac']))"; }
EXTRA="" run base A=1
EXTRA="" run s5 DMX_LIB=$P/libdmx_s5.so
EXTRA="" run s6 DMX_LIB=$P/libdmx_s6.so
EXTRA="--opt tile_qubits=7" run k7 A=1
EXTRA="--opt stream_chunk=32" run chunk32 A=1
EXTRA="--opt stream_chunk=128" run chunk128 A=1
echo done
