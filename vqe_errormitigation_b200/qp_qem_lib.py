"""Collaborator dialect of quasi-probability error mitigation (mirror of the reference's root-level
``QP_QEM_lib.py``): gate classes built from H and S, the 16 basis operations as gate tuples, PTM builders,
``lin_comb``, the SWAP-test circuit and ``QPSamp``.

Differences from the reference, on purpose:
  * ``QPSamp.sampled_measurement`` draws all ``sample_reps`` circuit variants at once and evaluates them as ONE GPU
    batch (the reference loops serially, "will be parallelized", QP_QEM_lib.py:753); shots are then drawn from
    the exact outcome probability of each variant;
  * ``sampled_measurement`` applies projector basis operations in their matrix form (``basis_ops``: ``Pz``).  The
    reference builds its sampled circuit from ``basis_ops_sim`` (QP_QEM_lib.py:787-791), whose projectors are
    ``MeasurementGate(1, key='M')`` — the key of the terminal measurement, so a circuit that drew one is rejected by
    cirq's ``run()`` ("Measurement key M repeated"); for the noise models this code can mitigate the
    quasi-probability of those operations is exactly zero, so they are never drawn.  The
    measurement dialect itself is available: ``sampled_circuit_sim`` builds that circuit with distinct keys and
    ``postselected_expectation`` runs it through the engine's mid-circuit measurement path (``trajectories``) with
    the post-selection Ref. 2 prescribes;
  * plotting at the end of ``run_experiment`` is omitted; undefined names of the original (``protocols`` at :135,
    ``error_type`` at :542) are not reproduced.
"""
from __future__ import annotations

import itertools as it
import math
from typing import List, Sequence

import numpy as np

from . import cirq_compat as cirq
from . import engine
from .error_mitigation import TwoQubitDepolarizingChannel

_I = np.eye(2, dtype=complex)
_X = np.array([[0, 1], [1, 0]], dtype=complex)
_Y = np.array([[0, -1j], [1j, 0]], dtype=complex)
_Z = np.array([[1, 0], [0, -1]], dtype=complex)


class _PowerGate(cirq.SingleQubitGate):
    _label = "?"

    def __init__(self, exponent=1):
        self.exponent = exponent

    def _base(self) -> np.ndarray:
        raise NotImplementedError

    def _unitary_(self):
        return np.linalg.matrix_power(self._base(), self.exponent)

    def _value_equality_values_(self):
        return self.exponent

    def __str__(self):
        return f"{self._label}**{self.exponent}"

    __repr__ = __str__


class PhaseS(_PowerGate):
    _label = "S"

    def _base(self):
        return (_I - 1j * _Z) / np.sqrt(2)


class Rz(_PowerGate):
    _label = "Rz"

    def _base(self):
        return np.linalg.matrix_power(PhaseS()._unitary_(), 3)


class Rx(_PowerGate):
    _label = "Rx"

    def _base(self):
        h = cirq.unitary(cirq.H)
        return h @ np.linalg.matrix_power(PhaseS()._unitary_(), 3) @ h


class Ry(_PowerGate):
    _label = "Ry"

    def _base(self):
        return Rz(3)._unitary_() @ Rx(1)._unitary_() @ Rz(1)._unitary_()


class Pz(cirq.SingleQubitGate):
    def _unitary_(self):
        return 0.5 * (_I + _Z)

    def _value_equality_values_(self):
        return ()

    def __str__(self):
        return "P₃"


def two_qubit_depolarize(p: float) -> TwoQubitDepolarizingChannel:
    return TwoQubitDepolarizingChannel(p)


def multi_kron(tup) -> np.ndarray:
    out = np.eye(1)
    for op in tup:
        out = np.kron(out, op)
    return out


def apply_to_qubit(op, qubit):
    if isinstance(op, (list, tuple, np.ndarray)):
        return type(op)(map(lambda g: g.on(qubit), op)) if not isinstance(op, np.ndarray) else np.array([g.on(qubit) for g in op])
    return op.on(qubit)


def _basis_tuples(projector):
    rx, ry, rz = Rx(1), Ry(1), Rz(1)
    rx2, rz2 = Rx(2), Rz(2)
    rx3, rz3 = Rx(3), Rz(3)
    ops = [(cirq.I,), (rx2,), (rz2, rx2), (rz2,), (rx,), (ry,), (rz,), (rz2, rx), (rz, rx, rz), (rz, rx2),
           (rz, rx, projector(), rx3, rz3), (rx3, projector(), rx), (projector(),),
           (rz, rx3, projector(), rx3, rz3), (rz2, rx3, projector(), rx), (rx2, projector())]
    return ops


def basis_ops():
    """16 basis operations as gate tuples (circuit order), projectors as ``Pz`` matrices (QP_QEM_lib.py:246-304)."""
    return _basis_tuples(Pz)


def basis_ops_sim():
    """Same, with projectors as ``MeasurementGate(1, key='M')`` (QP_QEM_lib.py:191-243): "postselection detailed in
    Ref. 2" — a shot counts only when every such measurement reads 0 (``QPSamp.postselected_expectation``)."""
    return _basis_tuples(lambda: cirq.MeasurementGate(1, key="M"))


def basis_matrices() -> List[np.ndarray]:
    """The 16 tuples of ``basis_ops`` multiplied out (first gate of a tuple acts first)."""
    mats = []
    for tup in basis_ops():
        m = np.eye(2, dtype=complex)
        for g in tup:
            m = cirq.unitary(g) @ m
        mats.append(m)
    return mats


def noise_model(dim, error_type, error):
    if error_type is None:
        return cirq.I
    if error_type == "depolarize":
        if not isinstance(error, (list, tuple, np.ndarray)):
            if dim == 1:
                return cirq.depolarize(error)
            if dim == 2:
                return two_qubit_depolarize(error)
            return None
        return cirq.asymmetric_depolarize(*error)
    if error_type == "amp_damp":
        return cirq.amplitude_damp(error)
    raise TypeError("Noise model not recognized")


def noise_model_channel(rho, error_type, error):
    """Dense depolarising channel on a matrix (QP_QEM_lib.py:345-385)."""
    if error_type != "depolarize":
        raise TypeError("Noise model not recognized")
    n = int(math.log(len(rho), 2))
    paulis = [multi_kron(p) for p in it.product([_I, _X, _Y, _Z], repeat=n)]
    if not isinstance(error, (list, tuple, np.ndarray)):
        coeff = error / (len(paulis) - 1) * np.ones(len(paulis))
    else:
        error = np.array(error, dtype=float)
        coeff = np.insert(error, 0, sum(error) / (len(paulis) - 1))
    if sum(coeff[1:]) > 1:
        raise TypeError("px+py+pz>1")
    out = (1 - len(paulis) * coeff[0]) * rho
    for c, p in zip(coeff, paulis):
        out = out + c * ((p @ rho) @ p)
    return out


def swap_circuit(n, qubits, decompose=True):
    """SWAP-test circuit on 2n+1 qubits (QP_QEM_lib.py:388-419)."""
    assert len(qubits) == 2 * n + 1
    circuit = cirq.Circuit()
    circuit.append([cirq.H(qubits[0]), cirq.H(qubits[1])])
    for j in range(2, n + 1):
        circuit.append(cirq.CNOT(qubits[1], qubits[j]))
    for j in range(1, n + 1):
        fredkin = cirq.CSWAP(qubits[0], qubits[j], qubits[j + n])
        circuit.append(fredkin._decompose_() if decompose else fredkin)
    circuit.append(cirq.H(qubits[0]))
    circuit.append(cirq.measure(qubits[0], key="M"))
    return circuit


def noisy_circuit(circuit, error_type=None, error=None):
    out = cirq.Circuit()
    for op in circuit.all_operations():
        out.append(op)
        if not isinstance(op.gate, cirq.MeasurementGate):
            out.append(noise_model(dim=len(op.qubits), error_type=error_type, error=error)(*op.qubits))
    return out


def _gate_matrix(gate) -> np.ndarray:
    """Unitary of an operation or of a tuple/list of operations applied in order (terminal measurements ignored)."""
    if isinstance(gate, (tuple, list)):
        m = None
        for g in gate:
            u = cirq.Circuit(g).unitary()
            m = u if m is None else u @ m
        return m
    return cirq.Circuit(gate).unitary()


def _o_tilde(gate, n_qubits, noisy, error_type, error):
    u = _gate_matrix(gate)
    d = 2 ** n_qubits
    paulis = [multi_kron(p) for p in it.product([_I, _X, _Y, _Z], repeat=n_qubits)]
    out = np.zeros((d * d, d * d), dtype=complex)
    for k, q in enumerate(paulis):
        rho = (u @ q) @ u.conj().T
        if noisy:
            rho = noise_model_channel(rho, error_type, error)
        rho = rho / d
        for j, obs in enumerate(paulis):
            out[j, k] = np.trace(obs @ rho)
    return out


def o_tilde_1q(gate, noisy, error_type, error):
    """PTM of a 1-qubit gate (or gate tuple), optionally followed by the dense noise channel (QP_QEM_lib.py:451-517)."""
    return _o_tilde(gate, 1, noisy, error_type, error)


def o_tilde_1q_tomo(gate, repetitions, noisy, error, error_type="depolarize"):
    """Gate-set-tomography estimate of a 1-qubit operation from sampled shots (QP_QEM_lib.py:520-590): entry [j][k] is the
    expectation of Q_j in {I, X, Y, Z} after preparing k in {|0>, |1>, |+>, |+i>} (``tomo_ops``), applying ``gate`` and rotating
    the measurement axis (``U_ops``); noise after the preparation and after the rotation, and after every gate of ``gate`` when
    ``noisy``.  An operation holding a ``MeasurementGate`` keyed 'M' (``basis_ops_sim``) is post-selected on M = 0:
    (N_0 - N_1) / repetitions, i.e. the unnormalised projector.  The reference body reads a module-level ``error_type`` that it
    never defines (a NameError whenever ``noisy``); here it is a trailing keyword.  The 16 circuits go through
    ``cirq.Simulator().run`` like the reference's - terminal-only ones as one evolution + shots, the post-selected ones as
    outcome-tree trajectories (``trajectories.TrajectoryProgram``)."""
    ops = list(gate) if isinstance(gate, (tuple, list)) else [gate]
    qubit = ops[0].qubits[0]
    noise = noise_model(dim=1, error_type=error_type if noisy else None, error=error)(qubit)
    body = [(g, noise) for g in ops] if noisy else ops
    tomo_ops = [cirq.I(qubit), cirq.X(qubit), cirq.ry(np.pi / 2)(qubit), cirq.rx(-np.pi / 2)(qubit)]
    u_ops = [cirq.I(qubit), cirq.H(qubit), cirq.rx(np.pi / 2)(qubit), cirq.I(qubit)]
    o_q = np.zeros((4, 4), dtype=complex)
    for j, k in it.product(range(4), repeat=2):
        circuit = cirq.Circuit(tomo_ops[k], noise, body, u_ops[j], noise)
        postselected = circuit.has_measurements()
        circuit.append(cirq.measure(qubit, key="M_post"))
        result = cirq.Simulator().run(program=circuit, repetitions=repetitions)
        if postselected:
            hist = result.multi_measurement_histogram(keys=("M", "M_post"))
            n_0, n_1 = hist[(0, 0)], hist[(0, 1)]
            o_q[j][k] = (n_0 + n_1) / repetitions if j == 0 else (n_0 - n_1) / repetitions
        elif j == 0:
            o_q[j][k] = 1
        else:
            o_q[j][k] = 2 * result.histogram(key="M_post")[0] / repetitions - 1
    return o_q


def o_tilde_2q(gate, noisy, error_type, error):
    return _o_tilde(gate, 2, noisy, error_type, error)


def lin_comb(op, basis):
    """Solve sum_i x_i vec_F(basis_i) = vec_F(op); complex coefficients are kept (QP_QEM_lib.py:659-679)."""
    a = np.array([b.flatten("F") for b in basis]).T
    return np.linalg.solve(a, op.flatten("F"))


class QPSamp:
    def __init__(self, circuit, error_type, error):
        self.B_ops_matrix = basis_ops()
        self.B_ops_circuit = basis_ops_sim()
        dummy = list(circuit.all_qubits())[0]
        self.B_bar_1q = [o_tilde_1q(apply_to_qubit(op, dummy), noisy=False, error_type=error_type, error=0).real for op in self.B_ops_matrix]
        self.B_bar_2q = [np.kron(b1, b2) for b1, b2 in it.product(self.B_bar_1q, repeat=2)]
        self.cost, self.qp_dists = [], []
        self.error_type, self.error = error_type, error
        self.circuit = circuit
        self.circuit_noisy = noisy_circuit(circuit, error_type=error_type, error=error)
        self.exp_ideal, self.exp_noisy, self.exp_sampled = [], [], []
        self._program = None

    def B_ops_circuit_2q(self, qubits, index):
        pairs = list(it.product(self.B_ops_circuit, repeat=2))[index]
        return apply_to_qubit(pairs[0], qubits[0]) + apply_to_qubit(pairs[1], qubits[1])

    def qp_costs(self):
        """QP distribution per operation *instance*, measurement included with cost 1 (QP_QEM_lib.py:711-746)."""
        self.cost, self.qp_dists = [], []
        cache = {}
        for op in self.circuit.all_operations():
            if isinstance(op.gate, cirq.MeasurementGate):
                self.qp_dists.append(np.eye(1, 16, 0)[0].astype(complex))
                self.cost.append(1)
                continue
            key = (op.gate, len(op.qubits))
            if key not in cache:
                if len(op.qubits) == 1:
                    o = o_tilde_1q(op, False, self.error_type, 0)
                    o_noise = o_tilde_1q(op, True, self.error_type, self.error)
                    basis = self.B_bar_1q
                else:
                    o = o_tilde_2q(op, False, self.error_type, 0)
                    o_noise = o_tilde_2q(op, True, self.error_type, self.error)
                    basis = self.B_bar_2q
                cache[key] = lin_comb(o.dot(np.linalg.inv(o_noise)), basis)
            qp = cache[key]
            self.qp_dists.append(qp)
            self.cost.append(float(np.sum(np.abs(qp))))

    def _variant_program(self) -> engine.CircuitProgram:
        if self._program is not None:
            return self._program
        qubits = sorted(self.circuit.all_qubits())
        pos = {q: i for i, q in enumerate(qubits)}
        table = basis_matrices()
        b = engine.ProgramBuilder(len(qubits))
        measurements = []
        slot = 0
        for op in self.circuit.all_operations():
            idx = [pos[q] for q in op.qubits]
            if isinstance(op.gate, cirq.MeasurementGate):
                measurements.append((op.gate.key, idx, op.gate.invert_mask))
                continue
            b.add_unitary(idx, cirq.unitary(op))
            noise = noise_model(dim=len(idx), error_type=self.error_type, error=self.error)
            kraus = cirq.channel(noise)
            b.add_kraus(idx, kraus)
            b.add_variant(idx, table, slot=slot)
            b.add_kraus(idx, kraus, if_slot=slot)
            slot += 1
        self._program = engine.CircuitProgram(b, qubits=qubits, measurements=measurements)
        return self._program

    def _p0(self, prog: engine.CircuitProgram, probs: np.ndarray) -> np.ndarray:
        measured = [m for m in prog.measurements if m[0] == "M"][0][1]
        states = np.arange(probs.shape[1])
        zero = np.ones(probs.shape[1], dtype=bool)
        for p in measured:
            zero &= ((states >> (prog.n_qubits - 1 - p)) & 1) == 0
        return probs[:, zero].sum(axis=1)

    def sampled_measurement(self, meas_reps, sample_reps):
        """Draw ``sample_reps`` variant circuits, evaluate them as one batch, sample ``meas_reps`` shots each;
        appends mean(sign * <Z>) to ``exp_sampled`` (QP_QEM_lib.py:748-804)."""
        gate_idx = [i for i, op in enumerate(self.circuit.all_operations()) if not isinstance(op.gate, cirq.MeasurementGate)]
        rows = np.zeros((sample_reps, len(gate_idx)), np.uint8)
        negative = np.zeros(sample_reps, dtype=np.int64)
        for s, i in enumerate(gate_idx):
            qp = np.asarray(self.qp_dists[i])
            prob = np.abs(qp) / self.cost[i]
            choice = np.random.choice(len(qp), size=sample_reps, p=prob / prob.sum())
            rows[:, s] = choice
            negative += (qp[choice].real < 0)
        signs = np.where(negative % 2 == 1, -1.0, 1.0)
        prog = self._variant_program()
        p0 = self._p0(prog, prog.probabilities(None, rows, renormalise=True))
        rng = np.random.default_rng(np.random.randint(0, 2 ** 31 - 1))
        hits = rng.binomial(meas_reps, np.clip(p0, 0.0, 1.0))
        self.exp_sampled.append(float(np.mean(signs * (2.0 * hits / meas_reps - 1.0))))

    def sampled_circuit_sim(self, row: Sequence[int]):
        """The sampled circuit of identifier ``row`` (one basis-operation index per gate) in the ``basis_ops_sim`` dialect, as
        QP_QEM_lib.py:758-795 assembles it: gate, noise, the basis operation as a gate tuple with ``MeasurementGate``s where it
        projects, noise again unless the operation is the identity.  The inserted measurements get their own keys
        (``P<gate>.<qubit>.<k>``; the reference re-uses 'M').  Returns (circuit, inserted keys)."""
        out = cirq.Circuit()
        keys: List[str] = []

        def place(tup, qubit, tag):
            for k, g in enumerate(tup):
                if isinstance(g, cirq.MeasurementGate):
                    keys.append(f"P{tag}.{k}")
                    out.append(cirq.MeasurementGate(1, key=keys[-1]).on(qubit))
                else:
                    out.append(g.on(qubit))

        s = 0
        for op in self.circuit.all_operations():
            out.append(op)
            if isinstance(op.gate, cirq.MeasurementGate):
                continue
            noise = noise_model(dim=len(op.qubits), error_type=self.error_type, error=self.error)
            out.append(noise(*op.qubits))
            idx = int(row[s])
            if len(op.qubits) == 1:
                place(self.B_ops_circuit[idx], op.qubits[0], f"{s}.0")
            else:
                place(self.B_ops_circuit[idx >> 4], op.qubits[0], f"{s}.0")
                place(self.B_ops_circuit[idx & 15], op.qubits[1], f"{s}.1")
            if idx != 0:
                out.append(noise(*op.qubits))
            s += 1
        return out, keys

    def postselected_expectation(self, row: Sequence[int], meas_reps: int, simulator=None) -> float:
        """<Z> of the terminal measurement over the shots whose inserted measurements all read 0, per shot TAKEN (the
        unnormalised estimate post-selection gives: rejected shots count as 0)."""
        circuit, keys = self.sampled_circuit_sim(row)
        sim = simulator if simulator is not None else cirq.DensityMatrixSimulator()
        result = sim.run(circuit, repetitions=meas_reps)
        keep = np.ones(meas_reps, dtype=bool)
        for k in keys:
            keep &= ~result.measurements[k].any(axis=1)
        final = result.measurements["M"][:, 0].astype(np.int64)
        return float(np.sum(keep * (1 - 2 * final)) / meas_reps)

    def _plain_measurement(self, circuit, meas_reps) -> float:
        result = cirq.DensityMatrixSimulator().run(circuit, repetitions=meas_reps)
        return 2 * result.histogram(key="M")[0] / meas_reps - 1

    def ideal_measurement(self, meas_reps):
        self.exp_ideal.append(self._plain_measurement(self.circuit, meas_reps))

    def noisy_measurement(self, meas_reps):
        self.exp_noisy.append(self._plain_measurement(self.circuit_noisy, meas_reps))

    def run_experiment(self, meas_reps, sample_reps, stat_reps, verbose: bool = True):
        """``stat_reps`` repetitions of (ideal, noisy, mitigated) estimates; returns the list of
        (ideal, noisy, prod(cost) * sampled) triples (the reference prints them and plots a histogram)."""
        self.qp_costs()
        results = []
        for _ in range(stat_reps):
            self.exp_ideal, self.exp_noisy, self.exp_sampled = [], [], []
            self.ideal_measurement(meas_reps)
            self.noisy_measurement(meas_reps)
            self.sampled_measurement(meas_reps, sample_reps)
            mitigated = float(np.prod(self.cost) * np.mean(self.exp_sampled))
            results.append((float(np.mean(self.exp_ideal)), float(np.mean(self.exp_noisy)), mitigated))
            if verbose:
                print(f"Ideal mean = {results[-1][0]}")
                print(f"Noisy error = {abs(results[-1][0] - results[-1][1])}")
                print(f"Sampled error = {abs(results[-1][0] - mitigated)}")
        return results
