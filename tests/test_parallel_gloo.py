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


def _mu_worker(rank, world, port, out_dir):
    """set_identifier_range -> get_mu_effective end to end on two ranks whose numpy RNGs DIFFER: rank 0's
    identifier table must win on every rank before the rows are sharded.  The GPU evaluation is replaced by a
    deterministic function of the row, so the test runs on CPU and any mix-up of rows between ranks changes the result."""
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vqe_errormitigation_b200 import cirq_compat as cirq
    from vqe_errormitigation_b200.data_containers.helper_interfaces.i_noise_wrapper import INoiseModel
    from vqe_errormitigation_b200.error_mitigation import IErrorMitigationManager
    q = cirq.LineQubit.range(2)
    circuit = cirq.Circuit([cirq.H(q[0]), cirq.CNOT(q[0], q[1]), cirq.rx(0.3)(q[1])])
    noise = INoiseModel([cirq.depolarize(0.05)], [], "dep")
    mgr = IErrorMitigationManager(circuit, noise, None)
    np.random.seed(1000 + 17 * rank)                  # deliberately different draws per rank
    mgr.set_identifier_range(400 + 50 * rank)         # ... and different table lengths
    local_rows = mgr._rows.copy()
    measure = lambda rows, counts: np.cos(rows.astype(float) @ np.arange(1, rows.shape[1] + 1)) + 0.01 * rows.sum(axis=1)
    mgr._measure_rows = measure
    values, mu = mgr.get_mu_effective(error_mitigation=True, density_representation=True)
    np.save(os.path.join(out_dir, f"mu{rank}.npy"), np.array([mu, len(values), mgr._rows.shape[0]]))
    np.save(os.path.join(out_dir, f"rows{rank}.npy"), mgr._rows)
    np.save(os.path.join(out_dir, f"local{rank}.npy"), local_rows)
    if rank == 0:
        signs, weights = mgr._signs_and_weights(mgr._rows)
        want = float(np.sum(signs * np.prod(weights) * measure(mgr._rows, mgr._counts) * mgr._counts) / np.sum(mgr._counts))
        np.save(os.path.join(out_dir, "want.npy"), np.array([want]))
    dist.destroy_process_group()


def test_get_mu_effective_world_size_2_consistent_table(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_mu_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    mu0, mu1 = np.load(tmp_path / "mu0.npy"), np.load(tmp_path / "mu1.npy")
    want = np.load(tmp_path / "want.npy")[0]
    assert np.array_equal(np.load(tmp_path / "rows0.npy"), np.load(tmp_path / "rows1.npy"))      # one table after the call
    assert not np.array_equal(np.load(tmp_path / "local0.npy"), np.load(tmp_path / "local1.npy"))  # the draws did differ
    assert mu0[0] == mu1[0] and abs(mu0[0] - want) < 1e-12 and mu0[1] == mu1[1] == 400


def test_broadcast_table_without_process_group_is_identity():
    from vqe_errormitigation_b200 import parallel
    rows, counts = np.arange(6, dtype=np.uint8).reshape(3, 2), np.array([1, 2, 3])
    r, c = parallel.broadcast_table(rows, counts)
    assert r is rows and c is counts
