"""Step time of the noisy-H2 sweep against the batch size (fixed launch cost vs per-circuit cost).  Run on the GPU box:
    python tools/batch_scaling.py"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import torch

from workloads import h2_program

prog, _ = h2_program(p1=1e-3, p2=1e-3)
flush = torch.empty(192 << 20, dtype=torch.uint8, device="cuda")
for B in (8, 1024, 8192, 32768, 65536, 131072, 262144, 1 << 20, 1 << 22):
    a = torch.linspace(-0.5, 0.5, B, dtype=torch.float64, device="cuda").reshape(-1, 1)
    out = torch.empty(B, dtype=torch.float64, device="cuda")
    for _ in range(5):
        prog.energies(a, out=out)
    warm = []
    for _ in range(50):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        prog.energies(a, out=out)
        e1.record()
        torch.cuda.synchronize()
        warm.append(e0.elapsed_time(e1) * 1e3)
    warm.sort()
    ts = []
    for _ in range(50):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        prog.energies(a, out=out)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    med = ts[len(ts) // 2]
    print(f"B={B:8d}  L2 flushed: median {med:8.2f} us  min {ts[0]:8.2f} us  {B / med * 1e6:.3e} circuits/s   warm L2: median {warm[len(warm) // 2]:8.2f} us")
