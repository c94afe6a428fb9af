"""Wall time of CircuitProgram.energies_host (page-locked buffers, zero-copy) against the batch size; also the bare C call.
    python tools/host_call_scaling.py"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np

from workloads import h2_program
from vqe_errormitigation_b200 import engine

prog, _ = h2_program(p1=1e-3, p2=1e-3)
for B in (8, 8192, 32768, 65536, 262144, 1 << 20):
    a = engine.pinned_empty((B, 1))
    a[...] = np.linspace(-0.5, 0.5, B).reshape(-1, 1)
    out = engine.pinned_empty(B)
    for _ in range(10):
        prog.energies_host(a, out=out)
    n = 300
    t0 = time.perf_counter()
    for _ in range(n):
        prog.energies_host(a, out=out)
    py = (time.perf_counter() - t0) / n * 1e6
    lib, h = prog._lib, prog._handle
    pa, po = a.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)
    t0 = time.perf_counter()
    for _ in range(n):
        lib.dmx_energy_batch_host(h, B, pa, None, po)
    raw = (time.perf_counter() - t0) / n * 1e6
    print(f"B={B:8d}  energies_host {py:8.2f} us   bare dmx_energy_batch_host {raw:8.2f} us   {B / raw * 1e6:.3e} circuits/s")
