"""End-to-end host-call study under torchrun (one rank per GPU): the noisy-H2 sweep through dmx_energy_batch_host with
page-locked buffers, for several in-flight depths and transfer modes, all ranks active and rank 0 alone.
    torchrun --nproc-per-node N tools/e2e_scaling.py [quick]
DMX_BENCH_NUMA=0 switches the NUMA-local core pinning of bench.pin_rank_to_cores off (even split of the allowed cores)."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
import torch
import torch.distributed as dist

import bench
from vqe_errormitigation_b200 import engine

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
pinned = bench.pin_rank_to_cores(local, world) if world > 1 else (None, None)
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rows = [None] * world
    dist.all_gather_object(rows, (rank, pinned, sorted(os.sched_getaffinity(0))[:3]))
    if rank == 0:
        print("pinning (rank, (cores, numa node), first cores):", rows, flush=True)
QUICK = len(sys.argv) > 1 and sys.argv[1] == "quick"
B, STEPS = 65536, 400


def measure(prog, depth, active=True):
    params = engine.pinned_empty((B, 1))
    params[:, 0] = np.linspace(-0.5, 0.5, B)
    res = [engine.pinned_empty(B) for _ in range(depth)]

    def step(i):
        prog.energies_host(params, None, batch=B, out=res[i % depth], nowait=depth > 1)
        if depth > 1 and (i + 1) % depth == 0:
            prog.host_sync()
    if active:
        for i in range(8):
            step(i)
        prog.host_sync()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    if active:
        for i in range(STEPS):
            step(i)
        prog.host_sync()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt if active else 0.0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) / STEPS * 1e6


for zero_copy, streams in (((1, 4), (0, 1)) if QUICK else ((1, 4), (1, 2), (1, 1), (0, 1))):
    bench.PLAN_OPTIONS.clear()
    bench.PLAN_OPTIONS["host_zero_copy"] = zero_copy
    bench.PLAN_OPTIONS["host_streams"] = streams
    wl = bench.workload_h2_sweep()
    prog = bench.get_prog(wl)
    for depth in ((1, 8, 16) if QUICK else (1, 2, 4, 8)):
        if zero_copy == 0 and depth > 1:
            continue
        all_us = measure(prog, depth)
        solo_us = measure(prog, depth, active=rank == 0)
        if rank == 0:
            print(f"zero_copy={zero_copy} streams={streams} depth={depth}: all {world} ranks {all_us:7.2f} us/step ({world * B / all_us * 1e6:.3e} circuits/s)   "
                  f"rank 0 alone {solo_us:7.2f} us/step ({B / solo_us * 1e6:.3e})", flush=True)
if world > 1:
    dist.destroy_process_group()
