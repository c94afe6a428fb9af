set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -q -x -k "h2 or host or async or batch or group or option or collection" 2>&1 | tail -3
python - <<'PY'
import subprocess, json, sys
def run(wl, opts, steps=50):
    cmd = [sys.executable, "bench.py", "--workload", wl, "--steps", str(steps), "--warmup", "5", "--no-cpu-baseline", "--no-also"] + sum([["--opt", o] for o in opts], [])
    r = subprocess.run(cmd, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        ov = d.get("overlapped", {}).get("value", 0)
        print(f"{wl:9s} {str(opts):22s} value {d['value']:.4g} us {1e3*d['ms_per_step']:.2f} overlapped {ov:.4g} e2e {d['e2e']['value']:.4g} sync {d['e2e'].get('synchronous',{}).get('value',0):.4g} parity {d['parity_max_abs_err']}", flush=True)
    except Exception as e:
        print(wl, opts, "FAILED", r.stdout[-300:], r.stderr[-1500:], flush=True)
for o in ([], ["frame32s=0"], ["frame32s=40"], ["frame32s=48"], ["frame32s=56"], ["frame32s=64"]):
    run("h2_sweep", o)
run("h2_bonds", [], 20)
run("h2_bonds", ["frame32s=0"], 20)
PY
python - <<'PY'
# host cost of one asynchronous call: raw ctypes vs the Python wrapper, enqueue time vs total
import time, ctypes as C, numpy as np
import bench
from vqe_errormitigation_b200 import engine
wl = bench.workload_h2_sweep(); prog = bench.get_prog(wl)
B = wl["batch"]; params = wl["make_inputs"](0)[0]
pp = engine.pinned_empty(params.shape); pp[...] = params
outs = [engine.pinned_empty(B) for _ in range(8)]
lib, h = prog._lib, prog._handle
pptr = C.c_void_p(pp.ctypes.data); optr = [C.c_void_p(o.ctypes.data) for o in outs]
for n_inflight in (8, 32):
    for name in ("raw", "wrapper"):
        for rep in range(2):
            prog.host_sync(); t0 = time.perf_counter(); N = 2000
            for i in range(N):
                if name == "raw": lib.dmx_energy_batch_host_async(h, B, pptr, None, optr[i % 8])
                else: prog.energies_host(pp, None, batch=B, out=outs[i % 8], nowait=True)
                if (i + 1) % n_inflight == 0: prog.host_sync()
            t1 = time.perf_counter(); prog.host_sync(); t2 = time.perf_counter()
        print(f"{name:8s} in_flight {n_inflight}: {1e6*(t2-t0)/N:.2f} us/step total ({B*N/(t2-t0):.3e}/s)")
# enqueue-only cost: queue 200 calls behind a long spin
import torch
for name in ("raw", "wrapper"):
    prog.host_sync(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(200):
        if name == "raw": lib.dmx_energy_batch_host_async(h, B, pptr, None, optr[i % 8])
        else: prog.energies_host(pp, None, batch=B, out=outs[i % 8], nowait=True)
    t1 = time.perf_counter(); prog.host_sync()
    print(f"{name:8s} enqueue only: {1e6*(t1-t0)/200:.2f} us/call")
PY
