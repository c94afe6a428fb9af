"""Where the time of IErrorMitigationManager.get_mu_effective_sampled goes (run on the GPU box):
sampler kernel, variant evaluation, reduction, and the whole call.   python tools/qem_breakdown.py"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
import torch

import bench
from vqe_errormitigation_b200 import engine, parallel

wl = bench.workload_h2_qem()
B = wl["batch"]
mgr = wl["manager"]
step = lambda i: mgr.get_mu_effective_sampled(B, seed=20260102 + i)
for i in range(3):
    step(i)
sampler, scale = mgr._sampler_cache
prog = mgr.variant_program()


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps, (time.perf_counter() - t0) * 1e6 / reps


variants, signs = sampler.draw(B, 1)
values = prog.energies(None, variants, batch=B)
print("sampler      gpu %.1f us  wall %.1f us" % timed(lambda: sampler.draw(B, 1)))
print("energies     gpu %.1f us  wall %.1f us" % timed(lambda: prog.energies(None, variants, batch=B)))
print("reduce       gpu %.1f us  wall %.1f us" % timed(lambda: parallel.sharded_mean(values, signs)))
print("whole call   gpu %.1f us  wall %.1f us" % timed(lambda: step(7)))
print("non-identity fraction of entries: %.5f, rows with any: %.4f" % (float((variants != 0).double().mean()), float((variants != 0).any(dim=1).double().mean())))
