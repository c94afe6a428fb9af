"""One 64-circuit chunk of a streaming workload (all its passes), three times: for ncu captures.
    python tools/profile_stream.py hea10|lih12 [key=value plan options ...]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch

import bench

name = sys.argv[1] if len(sys.argv) > 1 else "hea10"
for kv in sys.argv[2:]:
    k, v = kv.split("=")
    bench.PLAN_OPTIONS[k] = int(v)
wl = bench.WORKLOADS[name]()
prog = bench.get_prog(wl)
n = 64 if name == "hea10" else 8
params = torch.tensor(wl["make_inputs"](0)[0][:n], device="cuda")
for _ in range(3):
    e = prog.energies(params)
torch.cuda.synchronize()
print(name, prog.info["n_passes"], "passes", float(e[0]))
