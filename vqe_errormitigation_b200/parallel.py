"""Batch sharding across GPUs: one process per GPU, contiguous split of the circuit batch,
no data-path collective — only the per-circuit values (all-gather) or three partial sums (all-reduce, 24
bytes) cross NVLink.  Works with any initialised ``torch.distributed`` backend (nccl on the GPU box, gloo in
the CPU tests); without an initialised process group everything runs locally.

Replaces ``multiprocessing.Pool(6).map(process_circuit, ...)`` + ``np.mean`` of the reference
(``src/error_mitigation.py:419-427,451``).
"""
from __future__ import annotations

from typing import Callable, Tuple

import numpy as np


def _dist():
    try:
        import torch.distributed as dist
    except Exception:  # pragma: no cover
        return None
    return dist if dist.is_available() and dist.is_initialized() else None


def world() -> Tuple[int, int]:
    d = _dist()
    return (d.get_rank(), d.get_world_size()) if d else (0, 1)


def shard_bounds(total: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous split: rank g owns [g*total//G, (g+1)*total//G)."""
    return (rank * total) // world_size, ((rank + 1) * total) // world_size


def sharded_rows(fn: Callable[[np.ndarray, np.ndarray], np.ndarray], rows: np.ndarray, counts: np.ndarray) -> np.ndarray:
    """Evaluate ``fn(rows_shard, counts_shard)`` on this rank's shard and all-gather the per-row values."""
    d = _dist()
    if d is None:
        return np.asarray(fn(rows, counts), dtype=np.float64)
    import torch
    rank, size = d.get_rank(), d.get_world_size()
    lo, hi = shard_bounds(rows.shape[0], rank, size)
    local = np.asarray(fn(rows[lo:hi], counts[lo:hi]), dtype=np.float64) if hi > lo else np.zeros(0)
    device = torch.device("cuda", torch.cuda.current_device()) if d.get_backend() == "nccl" else torch.device("cpu")
    sizes = [shard_bounds(rows.shape[0], r, size) for r in range(size)]
    pad = max(h - l for l, h in sizes)
    buf = torch.zeros(pad, dtype=torch.float64, device=device)
    buf[:hi - lo] = torch.as_tensor(local, device=device)
    gathered = [torch.empty_like(buf) for _ in range(size)]
    d.all_gather(gathered, buf)
    return np.concatenate([g[:h - l].cpu().numpy() for g, (l, h) in zip(gathered, sizes)])


def broadcast_table(rows: np.ndarray, counts: np.ndarray, src: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """Make a (rows uint8 [R, S], counts int64 [R]) variant table identical on every rank: rank ``src``'s copy wins.
    The host-side sampler (``IErrorMitigationManager.set_identifier_range``) draws from the process-local numpy RNG, so
    without this step every rank would hold a different table - of a different length - and the sharded evaluation would
    combine values of unrelated rows.  The reference has a single dict in one parent process (error_mitigation.py:484-497)."""
    d = _dist()
    if d is None or d.get_world_size() == 1:
        return rows, counts
    import torch
    device = torch.device("cuda", torch.cuda.current_device()) if d.get_backend() == "nccl" else torch.device("cpu")
    shape = torch.tensor(list(rows.shape) if d.get_rank() == src else [0, 0], dtype=torch.int64, device=device)
    d.broadcast(shape, src=src)
    n_rows, n_cols = int(shape[0]), int(shape[1])
    if d.get_rank() == src:
        r_t = torch.as_tensor(np.ascontiguousarray(rows, np.uint8)).to(device)
        c_t = torch.as_tensor(np.ascontiguousarray(counts, np.int64)).to(device)
    else:
        r_t = torch.empty((n_rows, n_cols), dtype=torch.uint8, device=device)
        c_t = torch.empty(n_rows, dtype=torch.int64, device=device)
    if n_rows:
        d.broadcast(r_t, src=src)
        d.broadcast(c_t, src=src)
    return r_t.cpu().numpy(), c_t.cpu().numpy()


def sharded_sums(values, weights, counts=None):
    """Device tensor [3] = (sum w*v, sum (w*v)^2, N) over ALL ranks: dmx_qp_reduce locally, then one all-reduce of three
    doubles (NCCL: on the current stream, nothing is copied to the host - the caller decides when to synchronise)."""
    from . import engine
    local = engine.qp_reduce(values, weights, counts)
    d = _dist()
    if d is not None and d.get_world_size() > 1:
        if d.get_backend() != "nccl":
            host = local.cpu()
            d.all_reduce(host)
            local = host.to(local.device)
        else:
            d.all_reduce(local)
    return local


def sharded_mean(values, weights, counts=None) -> Tuple[float, float, float]:
    """(mean, variance, N) of w*v over all ranks from the three sums of :func:`sharded_sums` (one 24-byte D2H read)."""
    s0, s1, n = (float(x) for x in sharded_sums(values, weights, counts).cpu())
    mean = s0 / n if n else float("nan")
    var = s1 / n - mean * mean if n else float("nan")
    return mean, var, n
