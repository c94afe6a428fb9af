"""Ansatz interfaces (mirror of ``src/data_containers/helper_interfaces/i_wave_function.py``).

``IGeneralizedUCCSD`` turns the CCSD amplitudes of a molecule into the gate list of the UCCSD ansatz:
each kept fermionic excitation ``fo`` is Jordan-Wigner transformed (``fo - fo^dagger``) and every resulting
Pauli string becomes a basis change, a CNOT ladder, ``rz(2*sign*alpha_k)`` on the last qubit, the reverse
ladder and the inverse basis change (reference lines 158-189).  Only the *construction* happens here; the
circuit is evaluated by the CUDA engine.
"""
from __future__ import annotations

from abc import abstractmethod
from typing import Callable, Dict, Iterable, List, Sequence, Tuple

import numpy as np
import sympy

from ... import cirq_compat as cirq
from ... import openfermion_compat as of
from .i_parameter import IParameter


class IMolecule:
    def __init__(self, molecule_params: IParameter):
        self._params_molecule = molecule_params
        self.molecule = self._generate_molecule(p=molecule_params)

    def get_molecule_params(self) -> IParameter:
        return self._params_molecule

    molecule_parameters = property(get_molecule_params)

    @abstractmethod
    def _generate_molecule(self, p: IParameter) -> of.MolecularData:
        raise NotImplementedError

    def update_molecule(self, p: IParameter) -> of.MolecularData:
        """Reload the molecule for new parameters (the fermionic operators are NOT recomputed — reference
        behaviour, i_wave_function.py:31-35 vs :117)."""
        self._params_molecule = p
        self.molecule = self._generate_molecule(p=p)
        return self.molecule


class IWaveFunction(IMolecule):
    def __init__(self, operator_params: IParameter, molecule_params: IParameter):
        self._params_operator = operator_params
        IMolecule.__init__(self, molecule_params)
        self.qubits = self._generate_qubits()

    def get_operator_params(self) -> IParameter:
        return self._params_operator

    operator_parameters = property(get_operator_params)

    def params(self) -> Iterable[sympy.Symbol]:
        return [sympy.Symbol(name) for name in self._params_operator]

    @abstractmethod
    def _generate_qubits(self) -> Sequence[cirq.Qid]:
        raise NotImplementedError

    @abstractmethod
    def operations(self, qubits: Sequence[cirq.Qid]):
        raise NotImplementedError

    @abstractmethod
    def initial_state(self, qubits: Sequence[cirq.Qid]):
        yield []

    @abstractmethod
    def observable_measurement(self) -> Tuple[Callable, int]:
        raise NotImplementedError


def _numpy119_sign(value) -> float:
    """``np.sign`` of a complex number as numpy 1.19.1 (the reference's pin) computed it: the sign of the real
    part, or of the imaginary part when the real part is zero — returned as a real float.  numpy >= 2 returns
    ``z/|z|`` (``+-1j`` for the purely imaginary JW coefficients), which would make every rz angle imaginary
    (SURVEY.md §0 item 6)."""
    z = complex(value)
    if z.real != 0.0:
        return float(np.sign(z.real))
    return float(np.sign(z.imag))


class IGeneralizedUCCSD(IWaveFunction):
    def __init__(self, molecule_params: IParameter):
        IMolecule.__init__(self, molecule_params)
        self._fermionic_operators = self._get_fermionic_operators()
        count = int(0.5 * len(self._fermionic_operators))
        self._params_operator = IParameter({f"alpha{i}": 0.0 for i in range(count)})
        self.qubits = self._generate_qubits()

    def _get_hamiltonian(self) -> of.InteractionOperator:
        return self.molecule.get_molecular_hamiltonian()

    def _get_fermionic_operators(self) -> List[of.FermionOperator]:
        """normal_ordered(uccsd_generator(t1, t2)) as a list of single-term operators (reference :100-111)."""
        generator = of.uccsd_generator(self.molecule.ccsd_single_amps, self.molecule.ccsd_double_amps)
        return list(of.normal_ordered(generator))

    def _generate_qubits(self) -> Sequence[cirq.Qid]:
        """ceil(sqrt(n))^2 grid qubits (reference :126-130: 4 for H2, 16 for a 12-qubit molecule)."""
        side = int(np.ceil(np.sqrt(of.count_qubits(self._get_hamiltonian()))))
        return [cirq.GridQubit(r, c) for r in range(side) for c in range(side)]

    def operations(self, qubits: Sequence[cirq.Qid]):
        symbols = list(self.params())
        for k, excitation in enumerate(self._fermionic_operators[::2]):
            strings = list(of.jordan_wigner(excitation - of.hermitian_conjugated(excitation)))
            yield IGeneralizedUCCSD.qubit_to_gate_operators(qubit_operators=strings, qubits=qubits, par=symbols[k])

    @staticmethod
    def pauli_basis_map() -> Dict[str, Callable[[cirq.Qid, int], cirq.GateOperation]]:
        """X -> H, Y -> rx(-+pi/2), Z -> I (reference :142-156)."""
        return {"X": lambda q, inv: cirq.H.on(q),
                "Y": lambda q, inv: cirq.rx(-inv * np.pi / 2).on(q),
                "Z": lambda q, inv: cirq.I.on(q)}

    @staticmethod
    def qubit_to_gate_operators(qubit_operators: List[of.QubitOperator], qubits: Sequence[cirq.Qid], par: sympy.Symbol):
        to_z_basis = IGeneralizedUCCSD.pauli_basis_map()
        for q_op in qubit_operators:
            (string, coefficient), = q_op.terms.items()
            sign = _numpy119_sign(coefficient)
            support = [index for index, _ in string]
            for index, pauli in string:
                yield to_z_basis[pauli](qubits[index], 1)
            for a, b in zip(support[:-1], support[1:]):
                yield cirq.CNOT(qubits[a], qubits[b])
            yield cirq.rz(2 * sign * par).on(qubits[support[-1]])
            for a, b in reversed(list(zip(support[:-1], support[1:]))):
                yield cirq.CNOT(qubits[a], qubits[b])
            for index, pauli in string:
                yield to_z_basis[pauli](qubits[index], -1)

    @abstractmethod
    def initial_state(self, qubits: Sequence[cirq.Qid]):
        yield []
