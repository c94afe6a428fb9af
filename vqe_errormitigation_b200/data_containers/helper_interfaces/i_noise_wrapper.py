"""Noise model + noisy-ansatz wrapper (mirror of ``src/data_containers/helper_interfaces/i_noise_wrapper.py``)."""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import numpy as np

from ... import cirq_compat as cirq
from ... import openfermion_compat as of
from ...circuit_noise_extension import Noisify
from .i_parameter import IParameter
from .i_wave_function import IWaveFunction


class INoiseModel:
    def __init__(self, noise_gates_1q: List[cirq.Gate], noise_gates_2q: List[cirq.Gate], description: str):
        self._noise_gates_1q = noise_gates_1q
        self._noise_gates_2q = noise_gates_2q
        self._description = description

    def get_callable(self) -> Callable[[], List[cirq.Gate]]:
        return lambda: self._noise_gates_1q

    def get_description(self) -> str:
        return self._description

    def magic_check(self):
        for gate in self._noise_gates_1q:
            if gate.num_qubits() != 1:
                raise ValueError("Should be single qubit noise gate")
        for gate in self._noise_gates_2q:
            if gate.num_qubits() != 2:
                raise ValueError("Should be two qubit noise gate")

    def get_operators(self, *qubits: cirq.Qid) -> List[cirq.Operation]:
        """Channel list matching the arity of the gate that was just applied (reference :32-38)."""
        if len(qubits) == 1:
            return [gate.on(*qubits) for gate in self._noise_gates_1q]
        if len(qubits) == 2:
            return [gate.on(*qubits) for gate in self._noise_gates_2q]
        raise NotImplementedError("Can only construct single and two qubit noise operators")

    def get_effective_gate(self, gate: np.ndarray) -> np.ndarray:
        """Matrix-level noise for the PTM construction: rho -> sum_k p_k B_k rho B_k^dagger for each mixture
        channel of the matching arity, in order; channels without a mixture are silently skipped
        (reference :40-66, the swallowed TypeError at :50-51)."""
        rows, cols = gate.shape
        if rows == cols == 2:
            channels = self._noise_gates_1q
        elif rows == cols == 4:
            channels = self._noise_gates_2q
        else:
            raise NotImplementedError
        for noise_gate in channels:
            try:
                mix = cirq.mixture(noise_gate)
            except TypeError:
                continue
            if getattr(noise_gate, "_mixture_", None) is None and not cirq.has_unitary(noise_gate):
                continue
            acc = np.zeros(gate.shape, dtype=complex)
            for p, basis in mix:
                acc += p * (basis @ gate @ basis.conj().T)
            gate = acc
        return gate

    @staticmethod
    def empty() -> "INoiseModel":
        return INoiseModel(noise_gates_1q=[], noise_gates_2q=[], description="(p=0)")


class INoiseWrapper(IWaveFunction, cirq.NoiseModel):
    def __init__(self, w_class: IWaveFunction, noise_model: INoiseModel):
        self._ideal_wave_function = w_class
        self._noise_model = noise_model
        self._noise_channel = noise_model.get_operators
        # same IParameter objects as the wrapped ansatz (aliasing relied upon by the reference's drivers)
        IWaveFunction.__init__(self, w_class.operator_parameters, w_class.molecule_parameters)
        cirq.NoiseModel.__init__(self)

    def observable_measurement(self):
        return self._ideal_wave_function.observable_measurement()

    def sampled_energy_program(self, circuit: cirq.Circuit = None, noise_wrapper=None):
        """Batched sampled-energy program of the wrapped ansatz on the noisy circuit (symbols kept) under this wrapper's noise."""
        return self._ideal_wave_function.sampled_energy_program(self.get_noisy_circuit() if circuit is None else circuit,
                                                                self._noise_channel if noise_wrapper is None else noise_wrapper)

    def _generate_qubits(self) -> Sequence[cirq.Qid]:
        return self._ideal_wave_function._generate_qubits()

    def _generate_molecule(self, p: IParameter) -> of.MolecularData:
        return self._ideal_wave_function._generate_molecule(p=p)

    def operations(self, qubits: Sequence[cirq.Qid]):
        yield Noisify.wrap_noise_type(self._ideal_wave_function.operations(qubits=qubits), self._noise_channel)

    def initial_state(self, qubits: Sequence[cirq.Qid]):
        yield Noisify.wrap_noise_type(self._ideal_wave_function.initial_state(qubits=qubits), self._noise_channel)

    def noisy_operation(self, op):
        return [op] + list(self._noise_channel(*op.qubits))

    @property
    def noise_model(self) -> INoiseModel:
        return self._noise_model

    def get_clean_circuit(self) -> cirq.Circuit:
        from ...processors.processor_quantum import QPU
        circuit = QPU.get_initial_state_circuit(w=self._ideal_wave_function)
        circuit.append(self._ideal_wave_function.operations(self._ideal_wave_function.qubits))
        return circuit

    def get_noisy_circuit(self) -> cirq.Circuit:
        from ...processors.processor_quantum import QPU
        circuit = QPU.get_initial_state_circuit(w=self)
        circuit.append(self.operations(self._ideal_wave_function.qubits))
        return circuit
