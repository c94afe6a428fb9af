"""Where the isolated-launch time of the H2 sweep goes: the bench protocol (device spin, L2 flush, start event, ONE launch, end
event) applied to different batch sizes, with and without the flush, and to an empty launch.

    python tools/launch_floor.py [frame32s option values ...]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from vqe_errormitigation_b200 import engine  # noqa: E402


def timed(fn, flush, n=40):
    ts = []
    for i in range(n + 5):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(300_000)
        if flush is not None:
            flush.zero_()
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        if i >= 5:
            ts.append(1e3 * s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    opts = [int(v) for v in sys.argv[1:]] or [0, 1]
    flush = torch.empty(192 << 20, dtype=torch.uint8, device="cuda")
    tiny = torch.zeros(32, device="cuda")
    for fl in (flush, None):
        med, lo = timed(lambda: tiny.add_(1.0), fl)
        print(f"empty-ish launch (32-element add)  flush={'yes' if fl is not None else 'no '}: median {med:6.2f} us  min {lo:6.2f} us")
    wl = bench.workload_h2_sweep()
    for opt in opts:
        prog = engine.CircuitProgram(wl["builder"], qubits=wl["qubits"], measurements=wl["measurements"], hamiltonian=wl["terms"],
                                     options={"frame32s": opt})
        for B in (8, 1024, 8192, 32768, 65536, 131072, 262144, 1048576):
            p = torch.tensor(np.linspace(-0.5, 0.5, B).reshape(B, 1), device="cuda")
            out = torch.empty(B, dtype=torch.float64, device="cuda")
            row = []
            for fl in (flush, None):
                med, lo = timed(lambda: prog.energies(p, None, batch=B, out=out), fl)
                row.append(f"flush={'yes' if fl is not None else 'no '} median {med:7.2f} us min {lo:7.2f} us")
            print(f"frame32s={opt:<3d} batch {B:8d}: " + "   ".join(row), flush=True)


if __name__ == "__main__":
    main()
