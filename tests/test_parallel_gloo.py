"""N > 1 path on CPU: world_size-2 gloo process group exercises the batch sharding + gather / all-reduce logic
of vqe_errormitigation_b200.parallel (the GPU box runs the same code over NCCL)."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vqe_errormitigation_b200 import parallel
    rows = np.arange(23 * 3, dtype=np.uint8).reshape(23, 3)
    counts = np.arange(1, 24)
    seen = []

    def fn(r, c):
        seen.append(len(r))
        return r.sum(axis=1).astype(float) * 0.5 + c

    vals = parallel.sharded_rows(fn, rows, counts)
    lo, hi = parallel.shard_bounds(23, rank, world)
    assert seen == [hi - lo]
    np.save(os.path.join(out_dir, f"vals{rank}.npy"), vals)
    # empty shard on some ranks
    few = parallel.sharded_rows(lambda r, c: np.ones(len(r)), rows[:1], counts[:1])
    assert few.tolist() == [1.0]
    dist.destroy_process_group()


def test_sharded_rows_world_size_2(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    rows = np.arange(23 * 3, dtype=np.uint8).reshape(23, 3)
    want = rows.sum(axis=1).astype(float) * 0.5 + np.arange(1, 24)
    for r in range(2):
        assert np.array_equal(np.load(tmp_path / f"vals{r}.npy"), want)


def test_shard_bounds_cover_batch():
    from vqe_errormitigation_b200.parallel import shard_bounds
    for total in (0, 1, 7, 65536, 1000000):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
