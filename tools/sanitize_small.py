"""Small run of every n <= 4 device path (plain / zero-copy / DMA host call, variant rows incl. the worklist path, sampler,
reduction) - sized for a run under compute-sanitizer where that tool is available (it is closed on the round-1 pool):
    compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np
import torch

from oracle import dm_oracle as O
from test_gpu_parity import qp_noisy_variant_program
from workloads import h2_program, h2_terms
from vqe_errormitigation_b200 import engine

prog, _ = h2_program(p1=1e-3, p2=1e-3)
for B in (1, 7, 64, 1031):
    a = torch.linspace(-0.3, 0.3, B, dtype=torch.float64, device="cuda").reshape(-1, 1)
    e = prog.energies(a)
    pin, out = engine.pinned_empty((B, 1)), engine.pinned_empty(B)
    pin[...] = a.cpu().numpy()
    prog.energies_host(pin, out=out)                     # zero-copy
    assert np.array_equal(out, e.cpu().numpy())
    prog.energies_host(a.cpu().numpy())                  # DMA pipeline
vprog, gates = qp_noisy_variant_program([O.bit_flip(1e-3), O.depolarize(1e-3)], [O.two_qubit_depolarize(1e-3)], h2_terms())
rng = np.random.default_rng(1)
rows = np.zeros((515, 122), np.uint8)
rows[rng.integers(0, 515, 200), rng.integers(0, 122, 200)] = rng.integers(1, 4, 200)
rows[5, 3] = 7                                           # R-type correction: general path through the worklist
ev = vprog.energies(None, torch.tensor(rows, device="cuda"))
qps = [rng.normal(size=k) for k in (16, 256, 16, 2, 1, 256, 16)]
sampler = engine.QpSampler(qps)
for batch in (1, 33, 1000):
    v, s = sampler.draw(batch, 5)
red = engine.qp_reduce(ev, torch.ones_like(ev))
torch.cuda.synchronize()
print("sanitize_small ok", float(red[0]))
