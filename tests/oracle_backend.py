"""Test double: runs a CircuitProgram's op table through the C oracle (oracle/dm_oracle.c) instead of the GPU.

Used ONLY by CPU tests of the host logic (circuit construction, QP bookkeeping, optimiser drivers).  The product
has no such path: without this monkeypatch every CircuitProgram method raises on a machine without CUDA."""
from __future__ import annotations

import numpy as np

from oracle import c_oracle as CO
from oracle import dm_oracle as O
from vqe_errormitigation_b200 import engine


def _batch(self, params, variants, batch):
    sizes = set()
    p = v = None
    if params is not None and self.n_params:
        p = np.ascontiguousarray(np.asarray(params, np.float64).reshape(-1, self.n_params))
        sizes.add(p.shape[0])
    if variants is not None:
        v = np.ascontiguousarray(np.asarray(variants, np.uint8).reshape(-1, self.n_slots))
        sizes.add(v.shape[0])
    if batch is not None:
        sizes.add(batch)
    if not sizes:
        sizes = {1}
    assert len(sizes) == 1
    return p, v, sizes.pop()


def _energies_host(self, params=None, variants=None, batch=None, groups=None):
    p, v, B = _batch(self, params, variants, batch)
    if groups is None:
        return CO.energy_batch(self.n_qubits, self.ops, self.pool, self.hamiltonian, params=p, variants=v, batch=B,
                               n_params=self.n_params, n_slots=self.n_slots, threads=2)
    # coefficient groups: the oracle evaluates the rows of every group with that group's weights
    groups = np.asarray(groups).reshape(-1)
    x, z, _ = self.hamiltonian
    out = np.empty(B)
    for g in np.unique(groups):
        rows = np.flatnonzero(groups == g)
        out[rows] = CO.energy_batch(self.n_qubits, self.ops, self.pool, (x, z, self.coefficient_groups[g]),
                                    params=None if p is None else p[rows], variants=None if v is None else v[rows], batch=len(rows),
                                    n_params=self.n_params, n_slots=self.n_slots, threads=2)
    return out


def _density(self, params=None, variants=None, *, batch=None):
    p, v, B = _batch(self, params, variants, batch)
    return np.stack([CO.density(self.n_qubits, self.ops, self.pool, params=None if p is None else p[b],
                                variants=None if v is None else v[b]) for b in range(B)])


def _probabilities(self, params=None, variants=None, *, batch=None, renormalise=True, as_tensor=False):
    rho = _density(self, params, variants, batch=batch)
    return np.stack([O.probabilities(r, renormalise) for r in rho])


class _FakeTensor:
    def __init__(self, arr):
        self._a = arr

    def cpu(self):
        return self

    def numpy(self):
        return self._a


def _energies(self, params=None, variants=None, *, batch=None, out=None, workspace_states=None, groups=None, stream=None):
    return _FakeTensor(_energies_host(self, params, variants, batch, groups))


def _shot_expectations(probs, shots, masks, seed, first_row=0):
    """numpy Philox restatement of dmx_shot_expectations (tests/philox_ref.py) on a host array."""
    import philox_ref
    arr = probs._a if isinstance(probs, _FakeTensor) else np.asarray(probs)
    if not isinstance(shots, (int, np.integer)):
        shots = np.asarray(shots)
    return _FakeTensor(philox_ref.shot_expectations(arr, shots, [int(m) for m in masks], int(seed), int(first_row)))


def _sampled_energies(self, params, meas_reps, seed):
    """SampledEnergyProgram.energies without a device: same rows, same Philox streams, same weights."""
    M = len(self.variant_rows)
    if params is None:
        params = self.lifted[None, :] if self.prog.n_params else None
    B = 1 if params is None else len(params)
    p = None if params is None else np.repeat(np.asarray(params, np.float64), M, axis=0)
    v = np.tile(self.variant_rows, (B, 1))
    probs = self.prog.probabilities(p, v, batch=B * M, renormalise=True)
    expect = _shot_expectations(probs, meas_reps, self.masks, seed).numpy()
    w = np.tile(self.weights, (B, 1))
    return (expect * w).sum(axis=1).reshape(B, M).sum(axis=1) + self.identity


def install(monkeypatch):
    from vqe_errormitigation_b200.data_containers import model_hydrogen
    monkeypatch.setattr(engine, "shot_expectations", _shot_expectations)
    monkeypatch.setattr(model_hydrogen.SampledEnergyProgram, "energies", _sampled_energies)
    monkeypatch.setattr(engine.CircuitProgram, "energies_host", _energies_host)
    monkeypatch.setattr(engine.CircuitProgram, "energies", _energies)
    monkeypatch.setattr(engine.CircuitProgram, "density", _density)
    monkeypatch.setattr(engine.CircuitProgram, "probabilities", _probabilities)
