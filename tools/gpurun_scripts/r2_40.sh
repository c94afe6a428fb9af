set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "async or host" 2>&1 | tail -2
python - <<'PY'
import time, numpy as np
import bench
from vqe_errormitigation_b200 import engine
B = 65536
for opts in ({"host_h2d_dma": 0}, {"host_h2d_dma": 1}, {"host_h2d_dma": 1, "host_streams": 8}, {"host_h2d_dma": 1, "host_streams": 2}):
    bench.PLAN_OPTIONS.clear(); bench.PLAN_OPTIONS.update(opts)
    wl = bench.workload_h2_sweep(); prog = bench.get_prog(wl)
    params = engine.pinned_empty((B, 1)); params[:, 0] = np.linspace(-0.5, 0.5, B)
    want = prog.energies_host(params)
    for depth in (8, 16, 32):
        res = [engine.pinned_empty(B) for _ in range(depth)]
        def step(i):
            prog.energies_host(params, None, batch=B, out=res[i % depth], nowait=True)
            if (i + 1) % depth == 0: prog.host_sync()
        for i in range(2 * depth): step(i)
        prog.host_sync(); t0 = time.perf_counter(); N = 40 * depth
        for i in range(N): step(i)
        prog.host_sync(); dt = time.perf_counter() - t0
        ok = all(np.array_equal(r, want) for r in res)
        print(f"{opts} depth {depth}: {1e6*dt/N:.2f} us/step ({B*N/dt:.3e}/s) equal={ok}", flush=True)
PY
