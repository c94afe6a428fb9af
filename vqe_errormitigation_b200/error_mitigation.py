"""Quasi-probability error mitigation (mirror of ``src/error_mitigation.py``).

Host side: the 16-operation basis set, Pauli-transfer matrices, the quasi-probability (QP) solve, the
per-gate sampling of correction operations and the assembly of the corrected circuits.  Device side: all
sampled circuit variants are evaluated as ONE batch — the clean circuit is compiled once with a variant slot
after every gate (``DMX_OP_VARIANT``) and the ``uint8[B, n_gates]`` identifier table selects the correction
per circuit (the reference builds and simulates one cirq.Circuit per unique identifier in a
``multiprocessing.Pool(6)``, error_mitigation.py:416-427).
"""
from __future__ import annotations

import itertools as it
from typing import Any, Callable, Dict, Iterator, List, Optional, Sequence, Tuple, Union

import numpy as np

from . import cirq_compat as cirq
from . import engine
from .data_containers.helper_interfaces.i_noise_wrapper import INoiseModel
from .data_containers.helper_interfaces.i_wave_function import IGeneralizedUCCSD, IWaveFunction

MAX_MULTI_PROCESSING_COUNT = 6  # kept for API compatibility; the GPU batch replaces the process pool

_I = np.eye(2, dtype=complex)
_X = np.array([[0, 1], [1, 0]], dtype=complex)
_Y = np.array([[0, -1j], [1j, 0]], dtype=complex)
_Z = np.array([[1, 0], [0, -1]], dtype=complex)
_PAULI = (_I, _X, _Y, _Z)


def _word(*factors: np.ndarray) -> np.ndarray:
    """Matrix product of the factors, leftmost factor applied last (ordinary operator product)."""
    out = np.eye(2, dtype=complex)
    for f in factors:
        out = out @ f
    return out


class BasisUnitarySet:
    """The 16 single-qubit basis operations of Endo/Benjamin/Li-style QP mitigation, generated from the
    Hadamard H, the phase S = (I - iZ)/sqrt(2) and the projector pi_z = (I + Z)/2
    (reference error_mitigation.py:18-136; order of get_basis_unitary_set is part of the identifier format)."""

    @staticmethod
    def _h_unitary() -> np.ndarray:
        return cirq.unitary(cirq.H)

    @staticmethod
    def _s_unitary() -> np.ndarray:
        return (_I - 1j * _Z) / np.sqrt(2)

    @staticmethod
    def piz_unitary() -> np.ndarray:
        return 0.5 * (_I + _Z)

    @staticmethod
    def rz_unitary() -> np.ndarray:
        s = BasisUnitarySet._s_unitary()
        return _word(s, s, s)

    @staticmethod
    def rx_unitary() -> np.ndarray:
        h = BasisUnitarySet._h_unitary()
        return _word(h, BasisUnitarySet.rz_unitary(), h)

    @staticmethod
    def i_unitary() -> np.ndarray:
        return _I.copy()

    @staticmethod
    def sx_unitary() -> np.ndarray:
        rx = BasisUnitarySet.rx_unitary()
        return _word(rx, rx)

    @staticmethod
    def sz_unitary() -> np.ndarray:
        rz = BasisUnitarySet.rz_unitary()
        return _word(rz, rz)

    @staticmethod
    def sy_unitary() -> np.ndarray:
        return _word(BasisUnitarySet.sx_unitary(), BasisUnitarySet.sz_unitary())

    @staticmethod
    def ry_unitary() -> np.ndarray:
        rz, rx = BasisUnitarySet.rz_unitary(), BasisUnitarySet.rx_unitary()
        return _word(rz, rz, rz, rx, rz)

    @staticmethod
    def ryz_unitary() -> np.ndarray:
        rz, rx = BasisUnitarySet.rz_unitary(), BasisUnitarySet.rx_unitary()
        return _word(rx, rz, rz)

    @staticmethod
    def rzx_unitary() -> np.ndarray:
        rz, rx = BasisUnitarySet.rz_unitary(), BasisUnitarySet.rx_unitary()
        return _word(rz, rx, rz)

    @staticmethod
    def rxy_unitary() -> np.ndarray:
        rz, rx = BasisUnitarySet.rz_unitary(), BasisUnitarySet.rx_unitary()
        return _word(rx, rx, rz)

    @staticmethod
    def pix_unitary() -> np.ndarray:
        rz, rx, pz = BasisUnitarySet.rz_unitary(), BasisUnitarySet.rx_unitary(), BasisUnitarySet.piz_unitary()
        return _word(rz, rz, rz, rx, rx, rx, pz, rx, rz)

    @staticmethod
    def piy_unitary() -> np.ndarray:
        rx, pz = BasisUnitarySet.rx_unitary(), BasisUnitarySet.piz_unitary()
        return _word(rx, pz, rx, rx, rx)

    @staticmethod
    def piyz_unitary() -> np.ndarray:
        rz, rx, pz = BasisUnitarySet.rz_unitary(), BasisUnitarySet.rx_unitary(), BasisUnitarySet.piz_unitary()
        return _word(rz, rz, rz, rx, rx, rx, pz, rx, rx, rx, rz)

    @staticmethod
    def pizx_unitary() -> np.ndarray:
        rz = BasisUnitarySet.rz_unitary()
        return _word(BasisUnitarySet.piy_unitary(), rz, rz)

    @staticmethod
    def pixy_unitary() -> np.ndarray:
        rx = BasisUnitarySet.rx_unitary()
        return _word(BasisUnitarySet.piz_unitary(), rx, rx)

    _ORDER = ("i", "sx", "sy", "sz", "rx", "ry", "rz", "ryz", "rzx", "rxy", "pix", "piy", "piz", "piyz", "pizx", "pixy")

    @staticmethod
    def get_basis_unitary_set() -> Iterator[np.ndarray]:
        for name in BasisUnitarySet._ORDER:
            yield getattr(BasisUnitarySet, f"{name}_unitary")()

    @staticmethod
    def get_basis_set(dim: int) -> Iterator[np.ndarray]:
        """dim 2: the 16 operations; dim 4: the 256 Kronecker pairs, index 16*a + b."""
        singles = list(BasisUnitarySet.get_basis_unitary_set())
        if dim == 2:
            yield from singles
        elif dim == 4:
            for a, b in it.product(singles, repeat=2):
                yield np.kron(a, b)
        else:
            raise NotImplementedError

    @staticmethod
    def get_t_map(dim: int) -> np.ndarray:
        t_map = np.array([[1, 1, 1, 1], [0, 0, 1, 0], [0, 0, 0, 1], [1, -1, 0, 0]])
        if dim == 2:
            return t_map
        if dim == 4:
            return np.kron(t_map, t_map)
        raise NotImplementedError

    @staticmethod
    def get_observable_set(dim: int) -> Iterator[np.ndarray]:
        if dim == 2:
            yield from (p.copy() for p in _PAULI)
        elif dim == 4:
            for a, b in it.product(_PAULI, repeat=2):
                yield np.kron(a, b)
        else:
            raise NotImplementedError

    @staticmethod
    def get_initial_state_set(gate: np.ndarray) -> Iterator[np.ndarray]:
        d = gate.shape[0]
        if gate.shape != (d, d) or d not in (2, 4):
            raise NotImplementedError
        for observable in BasisUnitarySet.get_observable_set(dim=d):
            yield (gate @ observable @ gate.conj().T) / d

    @staticmethod
    def get_ptm_basis_set(dim: int) -> Iterator[np.ndarray]:
        singles = [BasisUnitarySet.pauli_transfer_matrix(u) for u in BasisUnitarySet.get_basis_unitary_set()]
        if dim == 2:
            yield from singles
        elif dim == 4:
            for a, b in it.product(singles, repeat=2):
                yield np.kron(a, b)
        else:
            raise NotImplementedError

    @staticmethod
    def pauli_transfer_matrix(gate: np.ndarray, noise_wrapper: Callable[[np.ndarray], np.ndarray] = None) -> np.ndarray:
        """PTM[j, k] = Tr(Q_j . N(U Q_k U^dagger / d)) with Q the Pauli (pair) basis (reference :224-259)."""
        d = gate.shape[0]
        observables = list(BasisUnitarySet.get_observable_set(dim=d))
        states = list(BasisUnitarySet.get_initial_state_set(gate=gate))
        if noise_wrapper is not None:
            states = [noise_wrapper(rho) for rho in states]
        ptm = np.zeros((d * d, d * d), dtype=complex)
        for j, obs in enumerate(observables):
            for k, rho in enumerate(states):
                ptm[j, k] = np.trace(obs @ rho)
        return ptm

    @staticmethod
    def get_quasiprobabilities(target: np.ndarray, basis_set: List[np.ndarray]) -> List[float]:
        """Solve sum_i q_i vec(B_i) = vec(target) (column-major vec), keep real parts (reference :262-274)."""
        if target.shape != basis_set[0].shape:
            raise ValueError(f"Target gate \n{target}\n and Basis set operation matrices should have the same shape.")
        columns = np.array([b.flatten("F") for b in basis_set]).T
        try:
            return [float(v.real) for v in np.linalg.solve(columns, target.flatten("F"))]
        except np.linalg.LinAlgError:
            print("WARNING: Not able to solve system of linear equations with ill defined basis.")
            fallback = [0.0] * len(basis_set)
            fallback[0] = 1
            return fallback


class IBasisGateSingle(cirq.SingleQubitGate):
    """Arbitrary 2x2 matrix (possibly a projector) applied as rho -> B rho B^dagger (reference :277-287)."""

    def __init__(self, unitary: np.ndarray, exponent=1):
        self._unitary = np.asarray(unitary)
        self._exponent = exponent

    def _unitary_(self):
        return np.linalg.matrix_power(self._unitary, self._exponent)

    def _value_equality_values_(self):
        return (self._unitary, self._exponent)

    def __str__(self):
        return "Su.Op."


class IBasisGateTwo(cirq.TwoQubitGate):
    def __init__(self, unitary: np.ndarray, exponent=1):
        self._unitary = np.asarray(unitary)
        self._exponent = exponent

    def _unitary_(self):
        return np.linalg.matrix_power(self._unitary, self._exponent)

    def _value_equality_values_(self):
        return (self._unitary, self._exponent)

    def __str__(self):
        return "Su.Op."


class QuasiCircuitIdentifier:
    """One sampled circuit variant: a basis-operation index per gate, the overall sign and the per-gate QP
    weights (reference :303-346)."""

    def __init__(self, circuit: cirq.Circuit, gate_lookup: Dict[cirq.Gate, List[float]]):
        self.identifiers_id, self.sign, self.identifiers_weight = QuasiCircuitIdentifier._get_identifiers(circuit, gate_lookup)

    @staticmethod
    def _get_identifiers(circuit: cirq.Circuit, gate_lookup: Dict[cirq.Gate, List[float]]) -> Tuple[List[int], int, List[float]]:
        ids, chosen, weights = [], [], []
        for op in circuit.all_operations():
            if isinstance(op.gate, cirq.MeasurementGate):
                weights.append(1)
                continue
            qp = gate_lookup[op.gate]
            basis_id, weight = QuasiCircuitIdentifier._get_identifier(quasi_probabilities=qp)
            ids.append(basis_id)
            chosen.append(qp[basis_id])
            weights.append(weight)
        return ids, np.sign(np.prod(chosen)), weights

    @staticmethod
    def _get_identifier(quasi_probabilities: List[float]) -> Tuple[int, float]:
        c_weight = sum(abs(p) for p in quasi_probabilities)
        index = np.random.choice(range(len(quasi_probabilities)), p=np.abs(quasi_probabilities) / c_weight)
        return index, c_weight

    @classmethod
    def from_arrays(cls, ids: Sequence[int], sign: float, weights: Sequence[float]) -> "QuasiCircuitIdentifier":
        obj = cls.__new__(cls)
        obj.identifiers_id, obj.sign, obj.identifiers_weight = list(int(i) for i in ids), sign, list(weights)
        return obj

    @property
    def identity(self):
        return [0] * len(self.identifiers_id)

    def __eq__(self, other):
        return hash(other) == hash(self)

    def __hash__(self):
        return hash(tuple(self.identifiers_id))


class IErrorMitigationManager:
    """Manages the family of sampled circuit variants and their (batched) evaluation."""

    def __init__(self, clean_circuit: cirq.Circuit, noise_model: INoiseModel, hamiltonian_objective: Union[np.ndarray, IWaveFunction, None]):
        self.circuit = clean_circuit
        self._noise_model = noise_model
        self._gate_lookup = IErrorMitigationManager.get_qp_lookup(clean_circuit, noise_model.get_effective_gate)
        self._objective_measure_cost = 1
        if isinstance(hamiltonian_objective, IWaveFunction):
            self._hamiltonian_objective = None
            self._wave_function = hamiltonian_objective
            _, self._objective_measure_cost = self._wave_function.observable_measurement()
        else:
            self._hamiltonian_objective = hamiltonian_objective
            self._wave_function = None
        self._error_mitigation = False
        self._density_representation = True
        self._meas_reps = 1
        self._identifier_count = 0
        # variant table (vectorised form of the reference's dict identifier -> count)
        self._gate_ops = [op for op in clean_circuit.all_operations() if not isinstance(op.gate, cirq.MeasurementGate)]
        self._n_measure_ops = sum(1 for op in clean_circuit.all_operations() if isinstance(op.gate, cirq.MeasurementGate))
        self._rows = np.zeros((0, len(self._gate_ops)), np.uint8)
        self._counts = np.zeros(0, np.int64)
        self._dict_cache: Optional[Dict[QuasiCircuitIdentifier, int]] = None
        self._programs: Dict[Tuple, Any] = {}

    # ---- QP tables ------------------------------------------------------------------------------------------
    @staticmethod
    def get_qp_lookup(clean_circuit: cirq.Circuit, noise_wrapper: Callable[[np.ndarray], np.ndarray]) -> Dict[cirq.Gate, List[float]]:
        """Per distinct gate: N^-1 = PTM_clean . inv(PTM_noisy) decomposed over the basis-op PTMs (reference :499-526)."""
        lookup: Dict[cirq.Gate, List[float]] = {}
        basis_cache: Dict[int, List[np.ndarray]] = {}
        for op in clean_circuit.all_operations():
            if isinstance(op.gate, cirq.MeasurementGate) or op.gate in lookup:
                continue
            clean_gate = cirq.unitary(op)
            clean_ptm = BasisUnitarySet.pauli_transfer_matrix(clean_gate)
            noisy_ptm = BasisUnitarySet.pauli_transfer_matrix(clean_gate, noise_wrapper=noise_wrapper)
            n_inv = clean_ptm.dot(np.linalg.inv(noisy_ptm))
            d = clean_gate.shape[0]
            if d not in basis_cache:
                basis_cache[d] = list(BasisUnitarySet.get_ptm_basis_set(dim=d))
            lookup[op.gate] = BasisUnitarySet.get_quasiprobabilities(target=n_inv, basis_set=basis_cache[d])
        return lookup

    # ---- identifier sampling ----------------------------------------------------------------------------------
    @property
    def _dict_identifiers(self) -> Dict[QuasiCircuitIdentifier, int]:
        if self._dict_cache is None:
            signs, weights = self._signs_and_weights(self._rows)
            self._dict_cache = {QuasiCircuitIdentifier.from_arrays(row, sg, weights): int(c)
                                for row, sg, c in zip(self._rows, signs, self._counts)}
        return self._dict_cache

    def _qp_matrix(self) -> List[np.ndarray]:
        # the lookup is fixed at construction; hashing 122 gate objects per call costs ~1 ms, so the list is built once
        if getattr(self, "_qp_cache", None) is None:
            self._qp_cache = [np.asarray(self._gate_lookup[op.gate], dtype=np.float64) for op in self._gate_ops]
        return self._qp_cache

    def _signs_and_weights(self, rows: np.ndarray) -> Tuple[np.ndarray, List[float]]:
        qps = self._qp_matrix()
        weights = [float(np.sum(np.abs(qp))) for qp in qps] + [1] * self._n_measure_ops
        if rows.shape[0] == 0:
            return np.zeros(0), weights
        negative = np.zeros(rows.shape[0], dtype=np.int64)
        zero = np.zeros(rows.shape[0], dtype=bool)
        for s, qp in enumerate(qps):
            chosen = qp[rows[:, s]]
            negative += chosen < 0
            zero |= chosen == 0
        signs = np.where(zero, 0.0, np.where(negative % 2 == 1, -1.0, 1.0))
        return signs, weights

    def set_identifier_range(self, count: int):
        """Draw ``count // objective_cost`` identifiers (reference :484-489) — vectorised per gate over all
        draws, then deduplicated into (unique rows, counts)."""
        count = int(count / self._objective_measure_cost)
        self._identifier_count = count
        self._dict_cache = None
        qps = self._qp_matrix()
        rows = np.zeros((count, len(qps)), np.uint8)
        for s, qp in enumerate(qps):
            prob = np.abs(qp) / np.sum(np.abs(qp))
            rows[:, s] = np.random.choice(len(qp), size=count, p=prob)
        if count:
            self._rows, self._counts = np.unique(rows, axis=0, return_counts=True)
        else:
            self._rows, self._counts = rows, np.zeros(0, np.int64)

    def add_circuit_identifier(self):
        ident = QuasiCircuitIdentifier(circuit=self.circuit, gate_lookup=self._gate_lookup)
        row = np.array(ident.identifiers_id, np.uint8)[None, :]
        match = np.flatnonzero((self._rows == row).all(axis=1)) if self._rows.shape[0] else []
        if len(match):
            self._counts[match[0]] += 1
        else:
            self._rows = np.concatenate([self._rows, row])
            self._counts = np.concatenate([self._counts, [1]])
        self._dict_cache = None

    def get_identifiers(self) -> Iterator[List[int]]:
        for row in self._rows:
            yield [int(v) for v in row]

    # ---- circuits -----------------------------------------------------------------------------------------------
    @staticmethod
    def get_updated_circuit(clean_circuit: cirq.Circuit, noise_wrapper: Callable, circuit_identifier: List[int], include_measurements: bool = False) -> cirq.Circuit:
        """gate, noise, basis op[id] (always), noise again iff id != 0 (reference :528-572)."""
        result = cirq.Circuit()
        bases = {2: list(BasisUnitarySet.get_basis_set(2)), 4: None}
        for i, op in enumerate(clean_circuit.all_operations()):
            if isinstance(op.gate, cirq.MeasurementGate):
                if include_measurements:
                    result.append(op)
                continue
            result.append(op)
            arity = len(op.qubits)
            if arity == 1:
                corrector = IBasisGateSingle(unitary=bases[2][circuit_identifier[i]]).on(*op.qubits)
            elif arity == 2:
                a, b = divmod(int(circuit_identifier[i]), 16)
                corrector = IBasisGateTwo(unitary=np.kron(bases[2][a], bases[2][b])).on(*op.qubits)
            else:
                raise NotImplementedError("Can only construct single and two qubit operators")
            noise_ops = noise_wrapper(*op.qubits)
            result.append(noise_ops)
            result.append(corrector)
            if circuit_identifier[i] != 0:
                result.append(noise_ops)
        return result

    def _observable_terms(self, n_qubits: int):
        obj = self._hamiltonian_objective
        if obj is None:
            # kron(eye(d/2), Z): Z on the LAST qubit (reference :381)
            return np.array([0], np.uint32), np.array([1], np.uint32), np.array([1.0])
        if hasattr(obj, "toarray"):
            return engine.pauli_decompose(obj.toarray())
        obj = np.asarray(obj)
        # ndarray objective: `*` is element-wise, so only diag(H) contributes to the trace (reference :382)
        return engine.pauli_decompose(np.diag(np.diagonal(obj)))

    def _variant_program(self, tail_ops=(), with_hamiltonian: bool = True, measure_all: bool = False):
        """Clean circuit + noise + one variant slot per gate (+ optional measurement tail), compiled once."""
        key = (tuple(tail_ops), with_hamiltonian, measure_all)
        if key in self._programs:
            return self._programs[key]
        b, qubits, measurements = self.variant_builder(tail_ops)
        ham = self._observable_terms(len(qubits)) if with_hamiltonian else None
        prog = engine.CircuitProgram(b, qubits=qubits, measurements=measurements, hamiltonian=ham)
        self._programs[key] = prog
        return prog

    def variant_builder(self, tail_ops=()):
        """The op table of the variant family (host only, no plan is created): gate, noise, ``DMX_OP_VARIANT`` slot,
        noise-if-slot-nonzero per gate of the clean circuit - the batched form of get_updated_circuit (reference :528-572).
        Returns (ProgramBuilder, qubits, measurements)."""
        qubits = sorted(self.circuit.all_qubits())
        pos = {q: i for i, q in enumerate(qubits)}
        basis = list(BasisUnitarySet.get_basis_unitary_set())
        b = engine.ProgramBuilder(len(qubits))
        measurements = []
        slot = 0
        for op in self.circuit.all_operations():
            idx = [pos[q] for q in op.qubits]
            if isinstance(op.gate, cirq.MeasurementGate):
                measurements.append((op.gate.key, idx, op.gate.invert_mask))
                continue
            b.add_unitary(idx, cirq.unitary(op))
            noise_ops = self._noise_model.get_operators(*op.qubits)
            for nop in noise_ops:
                b.add_kraus(idx, cirq.channel(nop))
            b.add_variant(idx, basis, slot=slot)
            for nop in noise_ops:
                b.add_kraus(idx, cirq.channel(nop), if_slot=slot)
            slot += 1
        for op in tail_ops:
            idx = [pos[q] for q in op.qubits]
            if isinstance(op.gate, cirq.MeasurementGate):
                measurements.append((op.gate.key, idx, op.gate.invert_mask))
            elif cirq.has_unitary(op.gate) and getattr(op.gate, "_mixture_", None) is None:
                b.add_unitary(idx, cirq.unitary(op))
            else:
                b.add_kraus(idx, cirq.channel(op))
        return b, qubits, measurements

    # ---- public batched surface (evaluate_variants) ---------------------------------------------------
    def variant_program(self):
        """The compiled plan of the variant family with the objective as its Hamiltonian (density representation)."""
        return self._variant_program()

    def quasi_probabilities(self) -> List[np.ndarray]:
        """Quasi-probability vector of every gate slot, in variant-table column order."""
        return self._qp_matrix()

    def observable_terms(self):
        """(xmask, zmask, coeff) of the objective the density path evaluates (reference :377-383)."""
        return self._observable_terms(len(sorted(self.circuit.all_qubits())))

    def total_cost(self) -> float:
        """prod_g sum_i |qp_g[i]|: the scalar every sampled value is multiplied by (reference :399, `np.prod(weights)`)."""
        if getattr(self, "_cost_cache", None) is None:
            self._cost_cache = float(np.prod([np.sum(np.abs(qp)) for qp in self._qp_matrix()]))
        return self._cost_cache

    def variant_signs(self, variants: np.ndarray) -> np.ndarray:
        """sign prod_g qp_g[variants[b, g]] per row (reference :321-323)."""
        return self._signs_and_weights(np.asarray(variants, np.uint8))[0]

    def evaluate_variants(self, variants, signs=None, counts=None):
        """Batched density-path evaluation of QP circuit variants: ``variants[B, S]`` (uint8 identifiers, one column per
        gate of the clean circuit), optional ``signs[B]`` (default: computed from the quasi-probabilities) and ``counts[B]``
        (multiplicity of each row; default 1).  Host arrays go through ``dmx_energy_batch_host`` (only the rows given are
        shipped - pass UNIQUE rows + counts, the reference's own identifier -> count dictionary, :492-497); CUDA tensors
        stay on the device (``dmx_energy_batch`` + ``dmx_qp_reduce`` + the cross-rank all-reduce of three doubles).
        Returns (values[B] = Tr(rho_b H) un-weighted, mu = cost * sum(sign*count*value) / sum(count))."""
        prog = self._variant_program()
        scale = self.total_cost()
        try:
            import torch
            on_device = isinstance(variants, torch.Tensor) and variants.is_cuda
        except ImportError:   # pragma: no cover
            on_device = False
        if on_device:
            from . import parallel
            B = variants.shape[0]
            values = prog.energies(None, variants, batch=B)
            if signs is None:
                signs = torch.as_tensor(self.variant_signs(variants.cpu().numpy()), device=variants.device)
            w = signs if counts is None else signs * counts.to(torch.float64)
            s0, _s1, n = (float(x) for x in parallel.sharded_sums(values, w, counts).cpu())
            return values, scale * s0 / n if n else float("nan")
        rows = np.ascontiguousarray(variants, np.uint8)
        values = prog.energies_host(None, rows)
        sg = self.variant_signs(rows) if signs is None else np.asarray(signs, np.float64)
        cn = np.ones(rows.shape[0]) if counts is None else np.asarray(counts, np.float64)
        total = float(np.sum(cn))
        return values, scale * float(np.dot(sg * cn, values)) / total if total else float("nan")

    # ---- evaluation -----------------------------------------------------------------------------------------------
    def _measure_rows(self, rows: np.ndarray, counts: np.ndarray) -> np.ndarray:
        """One measurement value per variant row (un-weighted, un-signed)."""
        if self._density_representation:
            prog = self._variant_program()
            return prog.energies_host(None, rows)
        reps = self._meas_reps * np.asarray(counts, dtype=np.int64)
        if self._wave_function is not None:
            return self._sampled_energy_rows(rows, reps)
        prog = self._variant_program(with_hamiltonian=False)
        if not prog.measurements:
            raise ValueError("Circuit has no measurements to sample.")
        key_pos = [m for m in prog.measurements if m[0] == "M"]
        if not key_pos:
            raise KeyError("M")
        n = prog.n_qubits
        measured = key_pos[0][1]
        if len(measured) == 1:
            # histogram('M')[0]: one measured qubit -> 2 p0 - 1 is the parity expectation of that qubit (reference :391-396);
            # shots drawn on the device
            probs = prog.probabilities(None, rows, renormalise=True, as_tensor=True)
            seed = int(np.random.randint(0, 2 ** 31 - 1))
            return engine.shot_expectations(probs, reps, [1 << (n - 1 - measured[0])], seed).cpu().numpy()[:, 0]
        probs = prog.probabilities(None, rows, renormalise=True)
        states = np.arange(probs.shape[1])
        all_zero = np.ones(probs.shape[1], dtype=bool)
        for pos in measured:
            all_zero &= ((states >> (n - 1 - pos)) & 1) == 0
        p0 = probs[:, all_zero].sum(axis=1)
        rng = np.random.default_rng(np.random.randint(0, 2 ** 31 - 1))
        hits = rng.binomial(reps, np.clip(p0, 0.0, 1.0))
        return 2.0 * hits / reps - 1.0

    def _sampled_energy_rows(self, rows: np.ndarray, reps: np.ndarray) -> np.ndarray:
        """``measure_observable`` (model_hydrogen.py:114-175) for every variant row: the five measurement
        circuits are the variant circuit plus basis-change gates (+ their noise); shots are drawn per row."""
        from .openfermion_compat import jordan_wigner
        w = self._wave_function
        hamiltonian = jordan_wigner(w.molecule.get_molecular_hamiltonian())
        qubits = w.qubits
        n = len(qubits)
        to_z = IGeneralizedUCCSD.pauli_basis_map()      # static, like model_hydrogen.py:122: the objective may be an INoiseWrapper
        noise = self._noise_model.get_operators
        by_size: Dict[int, list] = {}
        for term in hamiltonian.terms:
            by_size.setdefault(len(term), []).append(term)
        def tail(term_list):
            ops = []
            for term in term_list:
                for q, p in term:
                    ops.append(to_z[p](qubits[q], 1))
                    ops.extend(noise(qubits[q]))
            return tuple(ops)

        seed = int(np.random.randint(0, 2 ** 31 - 1))
        reps_t = np.asarray(reps, np.int64)
        bitmask = lambda qs: sum(1 << (n - 1 - q) for q in qs)
        stream_row = [0]

        def expectations(tail_ops, masks):
            """|diag rho| of every variant row with this measurement tail, shots and parities on the device
            (dmx_shot_expectations); only [rows, len(masks)] doubles come back."""
            prog = self._variant_program(tail_ops=tail_ops, with_hamiltonian=False)
            probs = prog.probabilities(None, rows, renormalise=True, as_tensor=True)
            out = engine.shot_expectations(probs, reps_t, masks, seed, first_row=stream_row[0]).cpu().numpy()
            stream_row[0] += rows.shape[0]
            return out

        total = np.full(rows.shape[0], 0.0)
        expectation = {(): np.ones(rows.shape[0])}
        z_terms = [((i, "Z"),) for i in range(n)] + [((i, "Z"), (j, "Z")) for i, j in it.combinations(range(n), 2)]
        z_exp = expectations(tail(by_size.get(1, [])), [bitmask([q for q, _ in t]) for t in z_terms])
        for k, t in enumerate(z_terms):
            expectation[t] = z_exp[:, k]
        for term in by_size.get(4, []):
            expectation[term] = expectations(tail([term]), [bitmask(range(n))])[:, 0]
        for term, coefficient in hamiltonian.terms.items():
            if term in expectation:
                total = total + complex(coefficient).real * expectation[term]
        return total

    def process_circuit(self, process_settings: Tuple[QuasiCircuitIdentifier, int]) -> List[float]:
        """One unique variant, ``count`` occurrences (reference :352-399) — evaluated by the same batched path."""
        ident, count = process_settings
        ids = ident.identifiers_id if self._error_mitigation else ident.identity
        sign = ident.sign if self._error_mitigation else 1
        weights = ident.identifiers_weight if self._error_mitigation else [1]
        rows = np.array([ids], np.uint8)
        value = self._measure_rows(rows, np.array([count]))[0]
        return [sign * np.prod(weights) * value] * count

    def get_mu_effective(self, error_mitigation: bool, density_representation: bool, meas_reps: int = 1) -> Tuple[np.ndarray, float]:
        """All unique variants as one GPU batch; returns (per-sample weighted values, their mean) like the
        reference (:401-452).  When torch.distributed is initialised the unique rows are sharded across ranks
        and the three partial sums are all-reduced (parallel.sharded_mean)."""
        self._error_mitigation = error_mitigation
        self._density_representation = density_representation
        self._meas_reps = meas_reps
        from . import parallel
        if error_mitigation:
            # rank 0's table wins: the identifiers were drawn from each process's own numpy RNG (set_identifier_range)
            self._rows, self._counts = parallel.broadcast_table(self._rows, self._counts)
            self._dict_cache = None
            rows, counts = self._rows, self._counts
            signs, weights = self._signs_and_weights(rows)
            scale = float(np.prod(weights))
        else:
            rows = np.zeros((1, len(self._gate_ops)), np.uint8)
            # one identity row with multiplicity count * meas_reps, exactly the reference's item (:434): process_circuit then
            # takes meas_reps * that many shots (meas_reps enters twice) and returns that many result entries - kept as is
            counts = np.array([self._identifier_count * meas_reps], np.int64)
            signs, scale = np.ones(1), 1.0
        if rows.shape[0] == 0:
            return np.zeros(0), float("nan")
        values = parallel.sharded_rows(lambda r, c: self._measure_rows(r, c), rows, counts)
        weighted = signs * scale * values
        results = np.repeat(weighted, counts)
        return results, float(np.sum(weighted * counts) / np.sum(counts))

    def sampled_sums(self, sampler, lo: int, hi: int, seed: int):
        """Device tensor [3] = (sum sign*value, sum value^2, N) over the samples [lo, hi) of Philox stream ``seed`` on this
        rank, all-reduced over the ranks of an initialised process group (NCCL: on the current stream, no host copy).
        Circuits that do not run on the warp-per-circuit frame kernels (the 7-qubit SWAP test: 128 KiB of state and ~80
        shared-memory sweeps per row) are deduplicated ON THE DEVICE first - the reference's identifier -> count dictionary
        (:492-497): at p = 1e-4 about 300 of 10 000 sampled rows are distinct.  torch.unique is plumbing (a device sort of
        the identifier rows); the evaluation of the distinct rows is the engine's launch."""
        import torch
        from . import parallel
        prog = self._variant_program()
        if prog.info["path"] == 1 and hi > lo and hasattr(sampler, "sampled_energies"):
            # register-resident plans: one kernel draws and evaluates (dmx_qp_sampled_energies), no identifier table
            values, signs = sampler.sampled_energies(prog, hi - lo, seed, first_sample=lo)
            return parallel.sharded_sums(values, signs)
        variants, signs = sampler.draw(hi - lo, seed, first_sample=lo)
        if prog.info["path"] != 1 and hi > lo:
            uniq, inverse, counts = _unique_rows(variants)
            signs_u = torch.zeros(uniq.shape[0], dtype=signs.dtype, device=signs.device).scatter_(0, inverse, signs)
            values = prog.energies(None, uniq, batch=uniq.shape[0])
            cf = counts.to(signs.dtype)
            local = torch.stack([torch.sum(cf * signs_u * values), torch.sum(cf * (signs_u * values) ** 2), torch.sum(cf)])
            d = parallel._dist()
            if d is not None and d.get_world_size() > 1:
                if d.get_backend() != "nccl":
                    host = local.cpu()
                    d.all_reduce(host)
                    local = host.to(local.device)
                else:
                    d.all_reduce(local)
            return local
        values = prog.energies(None, variants, batch=hi - lo)
        return parallel.sharded_sums(values, signs)

    def get_mu_effective_sampled(self, count: int, seed: int = 0) -> Tuple[float, float, int]:
        """Mitigated expectation from ``count`` variants drawn ON THE DEVICE (density representation): the sampler
        (engine.qp_sample_variants), the batched variant evaluation and the reduction never leave the GPU, so the only
        host traffic is the QP tables in and three doubles out.  Each rank of an initialised torch.distributed job draws
        its own slice [first, first + share) of one Philox stream; the estimator is all-reduced (24 bytes).
        Returns (mean, standard error, N).  Same estimator as get_mu_effective(True, True) (reference :401-452); on the
        frame kernels duplicates are simply evaluated again (cheaper than sorting them), larger circuits are deduplicated
        on the device (:meth:`sampled_sums`)."""
        from . import engine, parallel
        rank, world_size = parallel.world()
        lo, hi = parallel.shard_bounds(int(count), rank, world_size)
        if getattr(self, "_sampler_cache", None) is None:
            qps = self._qp_matrix()
            self._sampler_cache = (engine.QpSampler(qps), float(np.prod([np.sum(np.abs(qp)) for qp in qps])))
        sampler, scale = self._sampler_cache
        s0, s1, n = (float(x) for x in self.sampled_sums(sampler, lo, hi, seed).cpu())
        mean = s0 / n if n else float("nan")
        var = s1 / n - mean * mean if n else float("nan")
        mean, var = scale * mean, scale * scale * var
        return mean, float(np.sqrt(max(var, 0.0) / n)) if n else float("nan"), int(n)

    def __str__(self):
        return str(self._dict_identifiers)


_ROW_HASH = None


def _unique_rows(variants):
    """(distinct rows, inverse, counts) of a uint8 CUDA tensor [B, S].  torch.unique(dim=0) sorts the rows lexicographically
    (milliseconds for 10 000 x 77); a 64-bit multiplicative hash of every row and a 1-D unique of the keys is ~20x faster.
    The grouping is verified against the rows themselves (one boolean comes back) and falls back to the exact sort on a
    collision."""
    import torch
    global _ROW_HASH
    B, S = variants.shape
    if _ROW_HASH is None or _ROW_HASH.shape[0] < S or _ROW_HASH.device != variants.device:
        g = torch.Generator().manual_seed(0x5eed)
        _ROW_HASH = (torch.randint(-(2 ** 62), 2 ** 62, (max(S, 256),), generator=g, dtype=torch.int64) | 1).to(variants.device)
    keys = (variants.to(torch.int64) * _ROW_HASH[:S]).sum(dim=1)
    _uk, inverse, counts = torch.unique(keys, return_inverse=True, return_counts=True)
    first = torch.zeros(counts.shape[0], dtype=torch.int64, device=variants.device).scatter_(0, inverse, torch.arange(B, device=variants.device))
    uniq = variants[first]
    if not bool((uniq[inverse] == variants).all()):
        return torch.unique(variants, dim=0, return_inverse=True, return_counts=True)
    return uniq, inverse, counts


def reconstruct_from_basis(quasi_probabilities: List[float], basis_set: List[np.ndarray]) -> np.ndarray:
    """Inverse of the QP solve: sum_i q_i B_i (reference :578-582)."""
    if len(quasi_probabilities) != len(basis_set):
        raise ValueError("Quasi probabilities and basis set should have the same length")
    return sum(q * b for q, b in zip(quasi_probabilities, basis_set))


def simulate_error_mitigation(clean_circuit: cirq.Circuit, noise_model: INoiseModel, process_circuit_count: int, desc: str,
                              hamiltonian_objective: Union[np.ndarray, IWaveFunction, None] = None) -> Tuple[float, float, float]:
    """The reference's four-way comparison (:585-655): {noiseless, noisy} x {plain, mitigated} estimates of one circuit,
    returning (mu_ideal, mu_noisy, mu_effective).  The plain column is always read from the density matrix, the mitigated
    one from measurements unless the objective is an ndarray (:604,628).  Each cell is one variant batch on the GPU.  The
    2 x 2 histogram figure is drawn only when matplotlib is importable (the reference always draws; nothing here needs it)."""
    density = isinstance(hamiltonian_objective, np.ndarray)
    mu = {}
    cells = {}
    for i, noisy in enumerate((False, True)):
        for j, mitigate in enumerate((False, True)):
            manager = IErrorMitigationManager(clean_circuit=clean_circuit, noise_model=noise_model if noisy else INoiseModel.empty(),
                                              hamiltonian_objective=hamiltonian_objective)
            manager.set_identifier_range(process_circuit_count)
            values, mu[i, j] = manager.get_mu_effective(error_mitigation=mitigate, density_representation=True if j == 0 else density,
                                                        meas_reps=1)
            cells[i, j] = (values, manager)
    mu_ideal, mu_noisy, mu_effective = mu[0, 0], mu[1, 0], mu[1, 1]
    try:
        import matplotlib.pyplot as plt
    except ImportError:
        return mu_ideal, mu_noisy, mu_effective
    fig, axs = plt.subplots(2, 2, tight_layout=True)
    fig.suptitle(f'{desc} using {"density matrix" if density else "measurements"} (#mitigation circuits={process_circuit_count})', fontsize=16)
    for (i, j), (values, manager) in cells.items():
        axs[i][j].title.set_text(f"measurement histogram.\n(Info: {manager._noise_model.get_description()}, error mitigation={bool(j)})")
        axs[i][j].set_xlabel("relative measurement")
        axs[i][j].set_ylabel("Bin count")
        axs[i][j].hist(values, bins=20)
        if i:
            axs[i][j].axvline(x=mu[i, j], linewidth=1, color="r", label=f"C E[mu_eff] = {np.round(mu[i, j], 5)}\n|eps| = {np.round(abs(mu[i, j] - mu_ideal), 5)}")
        axs[i][j].axvline(x=mu_ideal, linewidth=1, color="g", label=f"mu_ideal = {np.round(mu_ideal, 5)}")
        axs[i][j].legend()
    return mu_ideal, mu_noisy, mu_effective


def get_hardcoded_noise(name: str, p: float) -> Callable[[np.ndarray], np.ndarray]:
    """Deprecated in the reference (:658-716), kept for callers: matrix-level noise G -> sum_k c_k P_k G P_k^dagger on a 2x2 or
    4x4 gate matrix.  'depolar': (1 - 4p/3) + p/3 per Pauli incl. I (4x4: 1 - 16p/15, p/15 over the 16 products);
    'dephase': (1 - p) + p Z G Z (4x4: (1 - p) + p/3 over IZ, ZI, ZZ); 'default': identity."""

    def conjugated(gate: np.ndarray, head: float, terms) -> np.ndarray:
        gate = np.asarray(gate)
        out = head * gate.astype(complex)
        for c, b in terms:
            out = out + c * (b @ (gate @ b.conj().T))
        return out

    def depolarize(gate: np.ndarray) -> np.ndarray:
        dim = np.asarray(gate).shape[0]
        if dim not in (2, 4):
            raise NotImplementedError
        paulis = list(BasisUnitarySet.get_observable_set(dim=dim))
        k = dim * dim
        return conjugated(gate, 1 - p * k / (k - 1), [(p / (k - 1), b) for b in paulis])

    def dephase(gate: np.ndarray) -> np.ndarray:
        dim = np.asarray(gate).shape[0]
        if dim == 2:
            return conjugated(gate, 1 - p, [(p, _Z)])
        if dim == 4:
            return conjugated(gate, 1 - p, [(p / 3, np.kron(a, b)) for a, b in ((_I, _Z), (_Z, _I), (_Z, _Z))])
        raise NotImplementedError

    table = {"default": lambda gate: gate, "depolar": depolarize, "dephase": dephase}
    if name not in table:
        raise NotImplementedError(name)
    return table[name]


def decomposition_time_test(gate: np.ndarray):
    """Times one quasi-probability decomposition and prints its reconstruction error (reference :728-738).  The reference
    solves the 2x2 / 4x4 ``gate`` against the raw operation matrices - a non-square system - and then compares shapes that
    differ; here the gate's Pauli-transfer matrix is decomposed in the PTM basis, the solve ``get_qp_lookup`` performs."""
    import time
    start = time.time()
    target = BasisUnitarySet.pauli_transfer_matrix(gate=gate)
    basis_set = list(BasisUnitarySet.get_ptm_basis_set(dim=gate.shape[0]))
    qp = BasisUnitarySet.get_quasiprobabilities(target=target, basis_set=basis_set)
    rebuilt = reconstruct_from_basis(qp, basis_set)
    print(f"Calculation time: {time.time() - start} sec.")
    print(f"Sum of quasi-probabilities: {sum(qp)}")
    print(f"Target vs Reconst. difference: {get_matrix_difference(rebuilt, target)}")


def get_matrix_difference(a: np.ndarray, b: np.ndarray) -> float:
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)))


# ---- channels (reference :743-910) -------------------------------------------------------------------------------


class TwoQubitDepolarizingChannel(cirq.TwoQubitGate):
    """(1 - 16p/15) on II plus p/15 on each of the 16 Pauli pairs *including* II (net identity weight 1 - p)."""

    def __init__(self, p: float) -> None:
        self._p = cirq.validate_probability(p / 15, "p")
        self._p_i = 1 - cirq.validate_probability(16 * p / 15, "sum_p")

    def _mixture_(self):
        terms = [(self._p_i, np.kron(_I, _I))]
        terms += [(self._p, np.kron(a, b)) for a, b in it.product(_PAULI, repeat=2)]
        return tuple(terms)

    def _has_mixture_(self) -> bool:
        return True

    def _value_equality_values_(self):
        return self._p

    @property
    def p_(self) -> float:
        return self._p

    def __repr__(self) -> str:
        return f"cirq.two_qubit_depolarize(p={self._p!r})"

    def __str__(self) -> str:
        return f"two_qubit_depolarize(p={self._p!r})"

    def _circuit_diagram_info_(self, args: "cirq.CircuitDiagramInfoArgs") -> "cirq.CircuitDiagramInfo":
        return cirq.protocols.CircuitDiagramInfo(wire_symbols=("D({:.3f})".format(self._p),) * 2)

    def _json_dict_(self) -> Dict[str, Any]:
        """cirq's ``obj_to_dict_helper(self, ['p'])`` (QP_QEM_lib.py:134-135) with the stored per-term probability."""
        return {"cirq_type": type(self).__name__, "p": self._p}


class SingleQubitPauliChannel(cirq.SingleQubitGate):
    def __init__(self, p_x: float, p_y: float, p_z: float) -> None:
        self._p_i = 1 - p_x - p_y - p_z
        self._p_x, self._p_y, self._p_z = p_x, p_y, p_z

    def _mixture_(self):
        return ((self._p_i, _I), (self._p_x, _X), (self._p_y, _Y), (self._p_z, _Z))

    @staticmethod
    def _has_mixture_() -> bool:
        return True

    def _value_equality_values_(self):
        return (self._p_x, self._p_y, self._p_z)

    def __repr__(self) -> str:
        return f"cirq.single_qubit_asymmetry_depolarize(p_x={self._p_x}, p_y={self._p_y}, p_z={self._p_z})"

    def __str__(self) -> str:
        return f"single_qubit_asymmetry_depolarize(p_x={self._p_x}, p_y={self._p_y}, p_z={self._p_z})"

    def _circuit_diagram_info_(self, args: "cirq.CircuitDiagramInfoArgs") -> "cirq.CircuitDiagramInfo":
        return cirq.protocols.CircuitDiagramInfo(wire_symbols=("D({:.3f})".format(self._p_i),))


class TwoQubitPauliChannel(cirq.TwoQubitGate):
    """Product of two single-qubit Pauli mixtures: 16 Kraus operators (reference :825-845)."""

    def __init__(self, p_x: float, p_y: float, p_z: float) -> None:
        self._p_i = 1 - p_x - p_y - p_z
        self._p_x, self._p_y, self._p_z = p_x, p_y, p_z

    def _single_mixture_(self):
        return ((self._p_i, _I), (self._p_x, _X), (self._p_y, _Y), (self._p_z, _Z))

    def _mixture_(self):
        single = self._single_mixture_()
        return tuple((pa * pb, np.kron(a, b)) for (pa, a), (pb, b) in it.product(single, repeat=2))

    @staticmethod
    def _has_mixture_() -> bool:
        return True

    def _value_equality_values_(self):
        return (self._p_x, self._p_y, self._p_z)

    def __repr__(self) -> str:
        return f"cirq.two_qubit_asymmetry_depolarize(p_x={self._p_x}, p_y={self._p_y}, p_z={self._p_z})"

    def __str__(self) -> str:
        return f"two_qubit_asymmetry_depolarize(p_x={self._p_x}, p_y={self._p_y}, p_z={self._p_z})"

    def _circuit_diagram_info_(self, args: "cirq.CircuitDiagramInfoArgs") -> "cirq.CircuitDiagramInfo":
        return cirq.protocols.CircuitDiagramInfo(wire_symbols=("D({:.3f})".format(self._p_i),) * 2)


_P0 = np.array([[1, 0], [0, 0]], dtype=complex)
_P1 = np.array([[0, 0], [0, 1]], dtype=complex)


class SingleQubitLeakageChannel(cirq.SingleQubitGate):
    """"Mixture" (1, |0><0|), (sqrt(1-p), |1><1|) exactly as the reference declares it (:860-871); unused there."""

    def __init__(self, p: float) -> None:
        self._p = p

    def _mixture_(self):
        return ((1, _P0), (np.sqrt(1 - self._p), _P1))

    @staticmethod
    def _has_mixture_() -> bool:
        return True

    def _value_equality_values_(self):
        return self._p

    def __str__(self) -> str:
        return f"single_qubit_leakage(p={self._p})"

    __repr__ = __str__

    def _circuit_diagram_info_(self, args: "cirq.CircuitDiagramInfoArgs") -> "cirq.CircuitDiagramInfo":
        return cirq.protocols.CircuitDiagramInfo(wire_symbols=("L({:.3f})".format(self._p),))


class TwoQubitLeakageChannel(cirq.TwoQubitGate):
    def __init__(self, p: float) -> None:
        self._p = p

    def _single_mixture_(self):
        return ((1, _P0), (np.sqrt(1 - self._p), _P1))

    def _mixture_(self):
        single = self._single_mixture_()
        return tuple((pa * pb, np.kron(a, b)) for (pa, a), (pb, b) in it.product(single, repeat=2))

    @staticmethod
    def _has_mixture_() -> bool:
        return True

    def _value_equality_values_(self):
        return self._p

    def __str__(self) -> str:
        return f"two_qubit_leakage(p={self._p})"

    __repr__ = __str__

    def _circuit_diagram_info_(self, args: "cirq.CircuitDiagramInfoArgs") -> "cirq.CircuitDiagramInfo":
        return cirq.protocols.CircuitDiagramInfo(wire_symbols=("L({:.3f})".format(self._p),) * 2)
