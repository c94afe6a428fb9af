"""Mid-circuit measurement: the per-shot half of ``cirq.DensityMatrixSimulator``.

cirq 0.8.2 evolves a circuit ONCE and samples the diagonal only when every measurement is terminal
(``DensityMatrixSimulator._run``: ``are_all_measurements_terminal`` -> ``_run_sweep_sample``).  Otherwise it simulates every
repetition on its own (``_run_sweep_repeat``), and at each ``MeasurementGate`` calls ``measure_density_matrix``: outcome
probabilities = |diag rho| summed over the other qubits and normalised to 1, one outcome drawn, rho projected onto it and
divided by that probability.  ``simulate`` does the same for its single run.  The reference reaches this path through
``basis_ops_sim`` (``QP_QEM_lib.py:191-243``: projector basis operations as ``MeasurementGate(1, key='M')`` inside the circuit).

Here the repetitions are not simulated one by one.  A collapse is a projector, and a projector is a one-qubit basis operation
the engine already batches (``DMX_OP_VARIANT``): every measured qubit becomes a variant slot over {not measured, |0><0|, |1><1|},
so "the state of a shot after its first j outcomes" is one ROW of a variant batch.  Sampling walks the outcome tree level by
level: the distinct outcome prefixes of all shots are evaluated together on the GPU by the program cut off at the next
measurement (``Tr(P_b rho)`` for every outcome ``b`` of the sites measured there = a grouped energy batch over Z-string
observables), every shot draws its next outcome from its prefix's probabilities, and so on; measurements with nothing after
them are drawn from the diagonal of the final states.  R shots over M measured bits therefore cost at most
``min(R, 2^j)`` rows at level j instead of R full simulations, the joint distribution of the records is the one cirq's loop
samples from, and — because each level conditions on the state *at* the measurement — circuits with non-trace-preserving
"gates" (the projector matrices of ``IBasisGateSingle``) follow cirq's renormalisation exactly.

``TrajectoryProgram.joint_distribution`` enumerates the whole tree; the parity tests compare it with the oracle's explicit
per-outcome projection of the complex density matrix (``oracle/dm_oracle.py::measurement_distribution``) to 1e-12.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import engine

MAX_GROUP_BITS = 6          # measured qubits evaluated by one grouped batch (2^g outcome rows per prefix)


class _Group:
    """Collapse sites measured back to back (nothing but other measurements between them), on distinct qubits."""

    def __init__(self, n_ops_before: int):
        self.n_ops_before = n_ops_before
        self.sites: List[Tuple[int, int]] = []       # (slot, qubit position)
        self.program: Optional[engine.CircuitProgram] = None


class TrajectoryProgram:
    """A resolved circuit with measurements anywhere, compiled for trajectory sampling."""

    def __init__(self, circuit, qubit_order=None, _lowered=None):
        sites: list = []
        if _lowered is not None:            # program_for_circuit has lowered the circuit already (for the cache key)
            builder, qubits, measurements, lifted, key, sites = _lowered
        else:
            builder, qubits, measurements, lifted, key = engine.lower_circuit(circuit, qubit_order, lift_numeric_rotations=True, sites=sites)
        keys = [m[0] for m in measurements]
        dup = sorted({k for k in keys if keys.count(k) > 1})
        if dup:
            raise ValueError("Measurement key {} repeated".format(",".join(dup)))        # cirq: _verify_unique_measurement_keys
        self.builder = builder
        self.qubits = qubits
        self.n_qubits = builder.n_qubits
        self.measurements = measurements             # (key, qubit positions, invert mask) in circuit order
        self.sites = sites                           # (slot, qubit position, n_ops before, measurement index, bit index)
        self.n_sites = len(sites)
        self.structure_key = key
        self.lifted = np.asarray(lifted, np.float64)
        # a site collapses the state mid-way when any non-measurement op follows it; the others see the final state
        is_site_op = np.zeros(len(builder.ops) + 1, bool)
        for s in sites:
            is_site_op[s[2]] = True
        last_gate = max((i for i in range(len(builder.ops)) if not is_site_op[i]), default=-1)
        self.collapse_sites = [s for s in sites if s[2] < last_gate]
        self.final_sites = [s for s in sites if s[2] >= last_gate]
        self.groups: List[_Group] = []
        for s in self.collapse_sites:
            g = self.groups[-1] if self.groups else None
            between = range(g.n_ops_before, s[2]) if g else ()
            if (g is None or len(g.sites) >= MAX_GROUP_BITS or any(not is_site_op[i] for i in between)
                    or any(s[1] == q for _, q in g.sites)):
                g = _Group(s[2])
                self.groups.append(g)
            g.sites.append((s[0], s[1]))
        self.program = engine.CircuitProgram(builder, qubits=qubits, measurements=measurements)

    # ---- programs cut off at a group of measurements -------------------------------------------------------
    def _group_program(self, g: _Group) -> engine.CircuitProgram:
        if g.program is None:
            n, k = self.n_qubits, len(g.sites)
            zmask = np.zeros(1 << k, np.uint32)
            for S in range(1 << k):
                for i, (_, q) in enumerate(g.sites):
                    if (S >> i) & 1:
                        zmask[S] |= 1 << (n - 1 - q)
            # Tr(P_b rho) = 2^-k sum_S (-1)^{b.S} Tr(Z_S rho): one coefficient row per outcome b of the group
            b, S = np.meshgrid(np.arange(1 << k), np.arange(1 << k), indexing="ij")
            par = np.zeros_like(b)
            for i in range(k):
                par ^= ((b & S) >> i) & 1
            rows = (1.0 - 2.0 * par) / (1 << k)
            ham = (np.zeros(1 << k, np.uint32), zmask, rows[0].copy())
            g.program = engine.CircuitProgram(self.builder.prefix(g.n_ops_before), qubits=self.qubits, hamiltonian=ham,
                                              coefficient_groups=rows)
        return g.program

    def _params(self, rows: int, params):
        if params is None:
            params = self.lifted
        if self.program.n_params == 0:
            return None
        return np.tile(np.asarray(params, np.float64).reshape(1, -1), (rows, 1))

    def _group_probabilities(self, g: _Group, codes: np.ndarray, params) -> np.ndarray:
        """``codes[U, M]`` (variant rows: 0 not measured, 1 / 2 collapsed onto 0 / 1) -> normalised outcome probabilities
        [U, 2^k] of the group's sites, as ``measure_density_matrix`` forms them: |.| of the diagonal sums, divided by their sum."""
        U, G = codes.shape[0], 1 << len(g.sites)
        if g.n_ops_before == 0:                    # measured before anything happened: |0..0>
            out = np.zeros((U, G))
            out[:, 0] = 1.0
            return out
        prog = self._group_program(g)
        v = np.repeat(codes, G, axis=0)
        grp = np.tile(np.arange(G, dtype=np.int32), U)
        e = prog.energies(self._params(U * G, params), v, groups=grp)
        p = np.abs(np.asarray(e.cpu().numpy(), np.float64).reshape(U, G))
        total = p.sum(axis=1, keepdims=True)          # a prefix of probability ~1e-300 leaves a numerically zero state: no outcomes
        return np.divide(p, total, out=np.zeros_like(p), where=total > 0.0)

    # ---- sampling ------------------------------------------------------------------------------------------
    def sample(self, repetitions: int, rng=np.random, params=None, collapse_final: bool = False):
        """Draw ``repetitions`` trajectories.  Returns ``(codes uint8 [R, M], weight float64 [R], final int64 [R] or None)``:
        the variant code of every site (1 + outcome; the sites with nothing after them stay 0 unless ``collapse_final``), the
        product of the normalised probabilities of the outcomes a shot collapsed onto (cirq divides rho by each of them), and
        the basis state drawn from the diagonal of the shot's final state when the circuit ends in measurements."""
        R, M = repetitions, self.n_sites
        codes = np.zeros((R, M), np.uint8)
        weight = np.ones(R)
        for g in self.groups:
            uniq, inv = np.unique(codes, axis=0, return_inverse=True)
            inv = inv.reshape(-1)
            probs = self._group_probabilities(g, uniq, params)
            drawn = _draw_rows(probs, inv, rng)
            weight *= probs[inv, drawn]
            for i, (slot, _) in enumerate(g.sites):
                codes[:, slot] = 1 + ((drawn >> i) & 1)
        final = None
        if self.final_sites:
            uniq, inv = np.unique(codes, axis=0, return_inverse=True)
            inv = inv.reshape(-1)
            probs = np.asarray(self.program.probabilities(self._params(len(uniq), params), uniq if M else None, batch=len(uniq),
                                                          renormalise=True))
            final = _draw_rows(probs, inv, rng)
            if collapse_final:
                # the collapse projects the MEASURED qubits only: the shot's weight is the marginal probability of their bits
                # (cirq divides by each terminal measurement's conditional probability in turn - the product is this marginal),
                # not the probability of the whole basis state that was drawn
                n = self.n_qubits
                measured = 0
                for slot, q, *_ in self.final_sites:
                    measured |= 1 << (n - 1 - q)
                    codes[:, slot] = 1 + ((final >> (n - 1 - q)) & 1)
                states = np.arange(probs.shape[1], dtype=np.int64)
                for lo in range(0, R, 4096):              # [shots, 2^n] agreement mask in bounded pieces
                    hi = min(R, lo + 4096)
                    agree = (states[None, :] & measured) == (final[lo:hi, None] & measured)
                    weight[lo:hi] *= (probs[inv[lo:hi]] * agree).sum(axis=1)
        return codes, weight, final

    def records(self, codes: np.ndarray, final: Optional[np.ndarray]) -> Dict[str, np.ndarray]:
        """Measurement records by key, ``[R, bits]`` int8 with the gate's invert mask applied (cirq: ``TrialResult.measurements``)."""
        n = self.n_qubits
        final_slots = {s[0] for s in self.final_sites}
        out = {}
        for mi, (key, positions, invert) in enumerate(self.measurements):
            bits = np.zeros((codes.shape[0], len(positions)), np.int8)
            for slot, q, _, m_index, bit in self.sites:
                if m_index != mi:
                    continue
                if slot in final_slots and final is not None and not codes[:, slot].any():
                    bits[:, bit] = (final >> (n - 1 - q)) & 1
                else:
                    bits[:, bit] = codes[:, slot].astype(np.int8) - 1
            for i, inv in enumerate(invert):
                if inv:
                    bits[:, i] ^= 1
            out[key] = bits
        return out

    def run(self, repetitions: int, rng=np.random, params=None) -> Dict[str, np.ndarray]:
        codes, _, final = self.sample(repetitions, rng, params)
        return self.records(codes, final)

    def simulate(self, rng=np.random, params=None):
        """One trajectory: ``(final_density_matrix, measurement records)``, every measurement collapsed and renormalised."""
        codes, weight, final = self.sample(1, rng, params, collapse_final=True)
        rho = self.program.density(self._params(1, params), codes if self.n_sites else None, batch=1)[0] / weight[0]
        return rho, {k: v[0] for k, v in self.records(codes, final).items()}

    def joint_distribution(self, params=None) -> Tuple[np.ndarray, np.ndarray]:
        """The whole outcome tree: ``(codes [L, M], probability [L])`` over every outcome string of the COLLAPSE sites with
        non-zero probability (final sites stay unmeasured) - what ``sample`` draws from, without the drawing."""
        codes = np.zeros((1, self.n_sites), np.uint8)
        prob = np.ones(1)
        for g in self.groups:
            p = self._group_probabilities(g, codes, params)
            G = p.shape[1]
            codes = np.repeat(codes, G, axis=0)
            drawn = np.tile(np.arange(G), len(prob))
            for i, (slot, _) in enumerate(g.sites):
                codes[:, slot] = 1 + ((drawn >> i) & 1)
            prob = (prob[:, None] * p).reshape(-1)
            keep = prob > 0.0
            codes, prob = codes[keep], prob[keep]
        return codes, prob

    def final_probabilities(self, codes: np.ndarray, params=None) -> np.ndarray:
        """Normalised diagonal of the final state of every outcome row ``codes[U, M]``."""
        return np.asarray(self.program.probabilities(self._params(len(codes), params), codes if self.n_sites else None,
                                                     batch=len(codes), renormalise=True))


def _draw_rows(probs: np.ndarray, row_of_shot: np.ndarray, rng) -> np.ndarray:
    """One index per shot from the distribution ``probs[row_of_shot[r]]`` (inverse CDF on one uniform per shot)."""
    cdf = np.cumsum(probs, axis=1)
    cdf[:, -1] = 1.0
    u = rng.random_sample(len(row_of_shot))
    return (u[:, None] >= cdf[row_of_shot]).sum(axis=1).astype(np.int64)


_CACHE: "OrderedDict[tuple, TrajectoryProgram]" = OrderedDict()
_CACHE_SIZE = 16


def program_for_circuit(resolved, qubit_order=None) -> Tuple[TrajectoryProgram, Optional[np.ndarray]]:
    """Structure-keyed cache (like ``engine.program_for_circuit``): circuits that differ only in rotation angles share the
    compiled prefix programs.  Returns the program and this circuit's angle row."""
    sites: list = []
    builder, qubits, measurements, lifted, key = engine.lower_circuit(resolved, qubit_order, lift_numeric_rotations=True, sites=sites)
    try:
        device = engine._require_cuda().cuda.current_device()
    except RuntimeError:
        device = -1
    full_key = (tuple(qubits), key, device)
    tp = _CACHE.get(full_key)
    if tp is None:
        tp = TrajectoryProgram(resolved, qubit_order, _lowered=(builder, qubits, measurements, lifted, key, sites))
        _CACHE[full_key] = tp
        while len(_CACHE) > _CACHE_SIZE:
            _CACHE.popitem(last=False)
    else:
        _CACHE.move_to_end(full_key)
    return tp, np.asarray(lifted, np.float64)
