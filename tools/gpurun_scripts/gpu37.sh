timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_sampler.py -q -m gpu -x -k "not streaming and not eleven and not ten_qubit and not twelve" 2>&1 | tail -3
run() { name=$1; shift; env "$@" python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-e2e $EXTRA > gpurun_out/x_$name.json 2>gpurun_out/x_$name.err; python -c "
import json; d=json.load(open('gpurun_out/x_$name.json')); print('$name value %.5g us %.2f'%(d['value'], d['ms_per_step']*1e3))"; }
EXTRA="" run base A=1
EXTRA="" run uncapped DMX_F32_UNCAPPED=1
EXTRA="" run base2 A=1
EXTRA="" run uncapped2 DMX_F32_UNCAPPED=1
EXTRA="--workload h2_qem" run qem_base A=1
EXTRA="--workload h2_qem" run qem_uncapped DMX_F32_UNCAPPED=1
A="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
ncu --set full --clock-control none --import-source on -k regex:dmx_frame32 -s 3 -c 1 -o gpurun_out/prof_frame32_v8 -f $A > gpurun_out/ncu_a.log 2>&1
echo done
