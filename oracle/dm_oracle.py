"""CPU oracle — TEST INFRASTRUCTURE ONLY (imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg; never by the product package).

A plain numpy restatement of the algorithm the reference's hot path delegates to
``cirq==0.8.2``'s ``DensityMatrixSimulator`` (pinned in ``requirements.txt:16``; not installable here, so it
is restated from its published behaviour — SURVEY.md §3.3, App. B.1):

  * state: tensor of shape (2,)*2n, first n axes = row (ket) bits, last n = column (bra) bits, big-endian,
    initial |0..0><0..0|;
  * unitary-like op U on qubits qs:   rho <- U rho U^dagger   (U is NOT validated: projector "gates" pass);
  * mixture / channel op {K_k}:       rho <- sum_k K_k rho K_k^dagger  (mixture (p,U) -> K = sqrt(p) U);
  * simulate() returns the final tensor reshaped to (2^n, 2^n);
  * run() with terminal measurements samples from |diag rho| / sum |diag rho|.

and of the observable evaluation at the reference's call sites:
  * ``(csc(rho) * H).diagonal().sum().real``  — ``src/processors/processor_quantum.py:36-39``
  * ``(H * rho).diagonal().sum().real`` with ndarray H = element-wise product —
    ``src/error_mitigation.py:377-383``.

PARITY PINNING: the reference has no tests and publishes no deterministic golden vector for a noisy energy.
This oracle is pinned against everything that does exist (tests/test_oracle.py, tests/test_persistence.py):
  * the 30 (hf_energy, fci_energy) pairs stored in the reference's ``src/mol_data/H2_*.hdf5`` (noiseless
    E(alpha=0) = HF, min_alpha E = FCI = lambda_min(H)) and the per-bond-length FCI values the reference wrote into
    ``src/classic_minimisation/H2_*semi_minimisation*``;
  * the reference-GENERATED noisy SWAP-test run ``src/classic_minimisation/SWAP_Similarity_P0.0001_R100_C10000``
    (100 x 10 000 shots, no optimiser in the loop): this oracle's exact noisy <Z0> = 0.4560239 sits inside the stored
    mean's standard error (0.456066 +- 0.00085) — a statistical pin of the noise model, the op order and cirq's
    Fredkin decomposition;
  * analytic quasi-probability values, trace / Hermiticity invariants, closed-form channel identities.
Beyond that — every other noisy energy and mitigated estimator to 1e-10 — **parity is unpinned by the reference**:
this restatement is the sole oracle there.

``dtype=np.complex64`` reproduces the precision the reference actually runs at (cirq's default).
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import numpy as np

# ---- op list format: ("u", qubits, U) | ("k", qubits, [K_0, K_1, ...]) | ("m", qubits, key) --------------
Op = Tuple[str, Tuple[int, ...], object]


def _apply_left(t: np.ndarray, m: np.ndarray, axes: Sequence[int]) -> np.ndarray:
    k = len(axes)
    m = m.reshape((2,) * (2 * k))
    out = np.tensordot(m, t, axes=(list(range(k, 2 * k)), list(axes)))
    return np.moveaxis(out, list(range(k)), list(axes))


def apply_op(t: np.ndarray, n: int, kind: str, qubits: Sequence[int], payload) -> np.ndarray:
    """cirq.protocols.apply_channel on the rank-2n tensor (src call sites: SURVEY.md §0 item 1)."""
    left = list(qubits)
    right = [q + n for q in qubits]
    if kind == "u":
        u = np.asarray(payload).astype(t.dtype)
        t = _apply_left(t, u, left)
        return _apply_left(t, np.conj(u), right)
    if kind == "k":
        acc = np.zeros_like(t)
        for kr in payload:
            kr = np.asarray(kr).astype(t.dtype)
            acc = acc + _apply_left(_apply_left(t, kr, left), np.conj(kr), right)
        return acc
    if kind == "m":
        return t
    raise ValueError(kind)


def simulate(n: int, ops: Iterable[Op], dtype=np.complex128) -> np.ndarray:
    """final_density_matrix of the op list applied to |0..0><0..0|."""
    t = np.zeros((2,) * (2 * n), dtype=dtype)
    t[(0,) * (2 * n)] = 1.0
    for kind, qubits, payload in ops:
        t = apply_op(t, n, kind, qubits, payload)
    return t.reshape(1 << n, 1 << n)


def energy_sparse(rho: np.ndarray, h) -> float:
    """Tr(rho H) as at processor_quantum.py:36-39 (H: scipy sparse or dense matrix)."""
    if hasattr(h, "toarray"):
        h = h.toarray()
    return float(np.trace(rho @ h).real)


def energy_elementwise(rho: np.ndarray, h: np.ndarray) -> float:
    """``(H * rho).diagonal().sum().real`` with ndarray H: element-wise product (error_mitigation.py:382)."""
    return float((np.asarray(h) * rho).diagonal().sum().real)


def probabilities(rho: np.ndarray, renormalise: bool = True) -> np.ndarray:
    """What DensityMatrixSimulator.run samples from: |diag rho| (/ its sum)."""
    p = np.abs(np.diagonal(rho)).astype(np.float64)
    return p / p.sum() if renormalise else p


# ---- measurement with collapse: the per-repetition path of DensityMatrixSimulator ------------------------
def _outcome_probabilities(t: np.ndarray, n: int, qubits: Sequence[int]) -> np.ndarray:
    """cirq/sim/density_matrix_utils.py::_probs [cirq-0.8.2]: |diag rho| summed over the unmeasured qubits for every
    outcome of ``qubits`` (big-endian over the listed qubits), then divided by its sum."""
    diag = np.abs(np.diagonal(t.reshape(1 << n, 1 << n))).astype(np.float64).reshape((2,) * n)
    k = len(qubits)
    probs = np.zeros(1 << k)
    for b in range(1 << k):
        sl = [slice(None)] * n
        for i, q in enumerate(qubits):
            sl[q] = (b >> (k - 1 - i)) & 1
        probs[b] = diag[tuple(sl)].sum()
    return probs / probs.sum()


def _collapse(t: np.ndarray, n: int, qubits: Sequence[int], outcome: int, prob: float) -> np.ndarray:
    """measure_density_matrix [cirq-0.8.2]: everything outside the outcome's row AND column slice is zeroed, then
    ``out /= probs[result]``."""
    k = len(qubits)
    mask = np.ones(t.shape, dtype=bool)
    sl = [slice(None)] * (2 * n)
    for i, q in enumerate(qubits):
        bit = (outcome >> (k - 1 - i)) & 1
        sl[q] = bit
        sl[q + n] = bit
    mask[tuple(sl)] = False
    out = t.copy()
    out[mask] = 0
    return out / prob


def run_repeat(n: int, ops: Iterable[Op], repetitions: int, rng=np.random, dtype=np.complex128):
    """DensityMatrixSimulator._run_sweep_repeat [cirq-0.8.2]: every repetition is simulated on its own and every "m" op
    draws an outcome and collapses the state.  Returns {key: uint8 [repetitions, bits]}."""
    ops = list(ops)
    records = {}
    for _ in range(repetitions):
        t = np.zeros((2,) * (2 * n), dtype=dtype)
        t[(0,) * (2 * n)] = 1.0
        for kind, qubits, payload in ops:
            if kind != "m":
                t = apply_op(t, n, kind, qubits, payload)
                continue
            probs = _outcome_probabilities(t, n, qubits)
            result = int(rng.choice(len(probs), p=probs))
            t = _collapse(t, n, qubits, result, probs[result])
            k = len(qubits)
            records.setdefault(payload, []).append([(result >> (k - 1 - i)) & 1 for i in range(k)])
    return {key: np.array(v, dtype=np.uint8) for key, v in records.items()}


def measurement_distribution(n: int, ops: Iterable[Op], dtype=np.complex128):
    """The distribution ``run_repeat`` samples from, by walking every branch: {tuple of all measured bits in circuit order:
    probability}, plus the collapsed final density matrix of every branch {bits: rho} (what ``simulate`` would return)."""
    ops = list(ops)
    dist, finals = {}, {}

    def walk(t, start, bits, prob):
        for i in range(start, len(ops)):
            kind, qubits, payload = ops[i]
            if kind != "m":
                t = apply_op(t, n, kind, qubits, payload)
                continue
            probs = _outcome_probabilities(t, n, qubits)
            k = len(qubits)
            for b in range(1 << k):
                if probs[b] > 0.0:
                    walk(_collapse(t, n, qubits, b, probs[b]), i + 1, bits + tuple((b >> (k - 1 - j)) & 1 for j in range(k)),
                         prob * probs[b])
            return
        dist[bits] = dist.get(bits, 0.0) + prob
        finals[bits] = t.reshape(1 << n, 1 << n)

    t0 = np.zeros((2,) * (2 * n), dtype=dtype)
    t0[(0,) * (2 * n)] = 1.0
    walk(t0, 0, (), 1.0)
    return dist, finals


# ---- gate / channel definitions (cirq 0.8.2 conventions, App. B.1) --------------------------------------

I2 = np.eye(2, dtype=complex)
X = np.array([[0, 1], [1, 0]], dtype=complex)
Y = np.array([[0, -1j], [1j, 0]], dtype=complex)
Z = np.array([[1, 0], [0, -1]], dtype=complex)
H = np.array([[1, 1], [1, -1]], dtype=complex) / np.sqrt(2)
CNOT = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=complex)
PAULIS = (I2, X, Y, Z)


def rx(t: float) -> np.ndarray:
    c, s = np.cos(t / 2), np.sin(t / 2)
    return np.array([[c, -1j * s], [-1j * s, c]], dtype=complex)


def ry(t: float) -> np.ndarray:
    c, s = np.cos(t / 2), np.sin(t / 2)
    return np.array([[c, -s], [s, c]], dtype=complex)


def rz(t: float) -> np.ndarray:
    return np.array([[np.exp(-0.5j * t), 0], [0, np.exp(0.5j * t)]], dtype=complex)


def mixture_to_kraus(mix: Sequence[Tuple[float, np.ndarray]]) -> List[np.ndarray]:
    return [np.sqrt(p) * np.asarray(u, dtype=complex) for p, u in mix]


def depolarize(p: float) -> List[np.ndarray]:
    return mixture_to_kraus([(1 - p, I2), (p / 3, X), (p / 3, Y), (p / 3, Z)])


def bit_flip(p: float) -> List[np.ndarray]:
    return mixture_to_kraus([(1 - p, I2), (p, X)])


def phase_flip(p: float) -> List[np.ndarray]:
    return mixture_to_kraus([(1 - p, I2), (p, Z)])


def asymmetric_depolarize(px: float, py: float, pz: float) -> List[np.ndarray]:
    """cirq.asymmetric_depolarize and SingleQubitPauliChannel (src/error_mitigation.py:797-810)."""
    return mixture_to_kraus([(1 - px - py - pz, I2), (px, X), (py, Y), (pz, Z)])


def two_qubit_pauli_channel(px: float, py: float, pz: float) -> List[np.ndarray]:
    """TwoQubitPauliChannel: product of two single-qubit Pauli mixtures (src/error_mitigation.py:825-845)."""
    single = [(1 - px - py - pz, I2), (px, X), (py, Y), (pz, Z)]
    return mixture_to_kraus([(pa * pb, np.kron(a, b)) for pa, a in single for pb, b in single])


def two_qubit_depolarize(p: float) -> List[np.ndarray]:
    """TwoQubitDepolarizingChannel: (1-16p/15) II plus p/15 on all 16 Pauli pairs incl. II
    (src/error_mitigation.py:749-766, QP_QEM_lib.py:87-108)."""
    mix = [(1 - 16 * p / 15, np.kron(I2, I2))]
    mix += [(p / 15, np.kron(a, b)) for a in PAULIS for b in PAULIS]
    return mixture_to_kraus(mix)


def amplitude_damp(g: float) -> List[np.ndarray]:
    return [np.array([[1, 0], [0, np.sqrt(1 - g)]], dtype=complex), np.array([[0, np.sqrt(g)], [0, 0]], dtype=complex)]


def phase_damp(g: float) -> List[np.ndarray]:
    return [np.array([[1, 0], [0, np.sqrt(1 - g)]], dtype=complex), np.array([[0, 0], [0, np.sqrt(g)]], dtype=complex)]


# ---- the 16 basis operations of BasisUnitarySet (src/error_mitigation.py:28-136) ---------


def basis_operations() -> List[np.ndarray]:
    s = (I2 - 1j * Z) / np.sqrt(2)
    mp = np.linalg.matrix_power
    r_z = mp(s, 3)
    r_x = H @ r_z @ H
    pi_z = 0.5 * (I2 + Z)
    sig_x = mp(r_x, 2)
    sig_z = mp(r_z, 2)
    sig_y = sig_x @ sig_z
    r_y = mp(r_z, 3) @ r_x @ r_z
    r_yz = r_x @ mp(r_z, 2)
    r_zx = r_z @ r_x @ r_z
    r_xy = mp(r_x, 2) @ r_z
    pi_x = mp(r_z, 3) @ mp(r_x, 3) @ pi_z @ r_x @ r_z
    pi_y = r_x @ pi_z @ mp(r_x, 3)
    pi_yz = mp(r_z, 3) @ mp(r_x, 3) @ pi_z @ mp(r_x, 3) @ r_z
    pi_zx = pi_y @ mp(r_z, 2)
    pi_xy = pi_z @ mp(r_x, 2)
    return [I2.copy(), sig_x, sig_y, sig_z, r_x, r_y, r_z, r_yz, r_zx, r_xy, pi_x, pi_y, pi_z, pi_yz, pi_zx, pi_xy]


def basis_operations_2q() -> List[np.ndarray]:
    b = basis_operations()
    return [np.kron(a, c) for a in b for c in b]


def pauli_transfer_matrix(gate: np.ndarray, noise=None) -> np.ndarray:
    """PTM[j,k] = Tr(Q_j N(U Q_k U^dagger / d)) (src/error_mitigation.py:224-259)."""
    d = gate.shape[0]
    obs = list(PAULIS) if d == 2 else [np.kron(a, b) for a in PAULIS for b in PAULIS]
    out = np.zeros((d * d, d * d), dtype=complex)
    for k, q in enumerate(obs):
        rho = gate @ q @ gate.conj().T / d
        if noise is not None:
            rho = noise(rho)
        for j, o in enumerate(obs):
            out[j, k] = np.trace(o @ rho)
    return out


def kraus_on_matrix(kraus: Sequence[np.ndarray]):
    return lambda rho: sum(k @ rho @ k.conj().T for k in kraus)


def quasi_probabilities(gate: np.ndarray, noise_kraus_list: Sequence[Sequence[np.ndarray]]) -> np.ndarray:
    """QP decomposition of N^-1 = PTM_clean inv(PTM_noisy) over the basis-op PTMs, column-major vec, real part
    (src/error_mitigation.py:262-274,516-524)."""
    def noise(rho):
        for kr in noise_kraus_list:
            rho = kraus_on_matrix(kr)(rho)
        return rho
    clean = pauli_transfer_matrix(gate)
    noisy = pauli_transfer_matrix(gate, noise)
    n_inv = clean @ np.linalg.inv(noisy)
    basis = basis_operations() if gate.shape[0] == 2 else basis_operations_2q()
    cols = np.array([pauli_transfer_matrix(b).flatten("F") for b in basis]).T
    return np.linalg.solve(cols, n_inv.flatten("F")).real


# ---- H2 ansatz, restated from SURVEY App. A.3 (i_wave_function.py:158-189, model_hydrogen.py:70-81) ---------

H2_STRINGS = (("XXYX", +1), ("YYYX", -1), ("YXXX", -1), ("XYXX", -1), ("YXYY", +1), ("XYYY", +1), ("XXXY", +1), ("YYXY", -1))


def h2_gate_list(alpha: float) -> List[Op]:
    """The 122 gates of the noiseless H2 UCCSD circuit: rx(pi) on q0,q1, then 8 Pauli-exponential blocks."""
    ops: List[Op] = [("u", (0,), rx(np.pi)), ("u", (1,), rx(np.pi))]
    for string, sign in H2_STRINGS:
        for q, p in enumerate(string):
            ops.append(("u", (q,), H if p == "X" else rx(-np.pi / 2)))
        for q in range(3):
            ops.append(("u", (q, q + 1), CNOT))
        ops.append(("u", (3,), rz(2.0 * sign * alpha)))
        for q in range(2, -1, -1):
            ops.append(("u", (q, q + 1), CNOT))
        for q, p in enumerate(string):
            ops.append(("u", (q,), H if p == "X" else rx(np.pi / 2)))
    return ops


def with_noise(gates: Sequence[Op], noise_1q: Sequence[Sequence[np.ndarray]], noise_2q: Sequence[Sequence[np.ndarray]]) -> List[Op]:
    """Noisify.wrap_noise_type (src/circuit_noise_extension.py:11-22): every gate followed by the channel list
    of its arity (INoiseModel.get_operators, i_noise_wrapper.py:32-38)."""
    out: List[Op] = []
    for kind, qubits, payload in gates:
        out.append((kind, qubits, payload))
        if kind == "m":
            continue
        for kr in (noise_1q if len(qubits) == 1 else noise_2q):
            out.append(("k", qubits, kr))
    return out


def variant_ops(gates: Sequence[Op], noise_1q, noise_2q, identifier: Sequence[int]) -> List[Op]:
    """IErrorMitigationManager.get_updated_circuit (src/error_mitigation.py:528-572): gate, noise,
    basis op[id] (always appended), noise again only when id != 0."""
    b1, b2 = basis_operations(), basis_operations_2q()
    out: List[Op] = []
    i = 0
    for kind, qubits, payload in gates:
        if kind == "m":
            continue
        noise = noise_1q if len(qubits) == 1 else noise_2q
        out.append((kind, qubits, payload))
        for kr in noise:
            out.append(("k", qubits, kr))
        ident = int(identifier[i])
        out.append(("u", qubits, (b1 if len(qubits) == 1 else b2)[ident]))
        if ident != 0:
            for kr in noise:
                out.append(("k", qubits, kr))
        i += 1
    return out


def pauli_terms_to_matrix(n: int, xmask: Sequence[int], zmask: Sequence[int], coeff: Sequence[float]) -> np.ndarray:
    """Dense H from (xmask, zmask, coeff) terms; qubit q <-> bit n-1-q; Y where both masks are set."""
    dim = 1 << n
    out = np.zeros((dim, dim), dtype=complex)
    for x, z, c in zip(xmask, zmask, coeff):
        m = np.eye(1, dtype=complex)
        for q in range(n):
            bit = 1 << (n - 1 - q)
            xb, zb = bool(int(x) & bit), bool(int(z) & bit)
            m = np.kron(m, Y if (xb and zb) else X if xb else Z if zb else I2)
        out += c * m
    return out


def from_circuit(circuit, qubit_order=None) -> Tuple[int, List[Op]]:
    """Adapter: any object with cirq's protocol surface (all_operations, op.gate, op.qubits, _unitary_ /
    _mixture_ / _channel_) -> oracle op list.  Resolution order follows cirq's apply_channel: unitary first,
    then mixture, then channel."""
    qubits = list(qubit_order) if qubit_order is not None else sorted(circuit.all_qubits())
    pos = {q: i for i, q in enumerate(qubits)}
    ops: List[Op] = []
    for op in circuit.all_operations():
        idx = tuple(pos[q] for q in op.qubits)
        gate = op.gate
        if type(gate).__name__ == "MeasurementGate":
            ops.append(("m", idx, getattr(gate, "key", "")))
            continue
        u = gate._unitary_() if hasattr(gate, "_unitary_") else None
        if u is not None and u is not NotImplemented:
            ops.append(("u", idx, np.asarray(u)))
        elif hasattr(gate, "_mixture_"):
            ops.append(("k", idx, mixture_to_kraus(gate._mixture_())))
        elif hasattr(gate, "_channel_"):
            ops.append(("k", idx, [np.asarray(k) for k in gate._channel_()]))
        else:
            raise TypeError(f"cannot simulate {gate!r}")
    return len(qubits), ops
