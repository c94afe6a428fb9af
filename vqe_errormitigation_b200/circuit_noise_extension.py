"""Noise insertion on operation trees (mirror of ``src/circuit_noise_extension.py``)."""
from __future__ import annotations

from random import uniform
from typing import Callable, List

import numpy as np

from . import cirq_compat as cirq


class Noisify:
    @staticmethod
    def wrap_noise_type(op_tree, noise_wrapper: Callable[..., List[cirq.Operation]]):
        """Every gate operation followed by ``noise_wrapper(*gate_op.qubits)`` (reference :11-22).  This fixes
        the op order of the noisy circuit: 122 gates -> 244 ops for H2 with one channel per arity."""
        out = []
        for op_list in op_tree:
            for gate_op in cirq.flatten_op_tree(op_list):
                out.append(gate_op)
                out.extend(noise_wrapper(*gate_op.qubits))
        return out

    @staticmethod
    def introduce_noise_type(op_tree, noise_func: Callable[[], List[cirq.Gate]]):
        """Legacy form (reference :24-42): after each operation, the gates drawn from ``noise_func`` applied to
        each of its qubits (what cirq.ConstantQubitNoiseModel.noisy_operation appends)."""
        out = []
        for op_list in op_tree:
            for gate_op in cirq.flatten_op_tree(op_list):
                out.append(gate_op)
                for noise_gate in noise_func():
                    out.extend(noise_gate.on(q) for q in gate_op.qubits)
        return out

    @staticmethod
    def introduce_noise_tree(op_tree):
        return Noisify.introduce_noise_type(op_tree=op_tree, noise_func=lambda: Noisify.uniform_rotation_gate(1000))

    @staticmethod
    def introduce_noise(circuit: cirq.Circuit, noise_func: Callable[[], List[cirq.Gate]]) -> cirq.Circuit:
        return cirq.Circuit(Noisify.introduce_noise_type(circuit.moments, noise_func), strategy=cirq.InsertStrategy.EARLIEST)

    @staticmethod
    def uniform_rotation_gate(sample_count: int) -> List[cirq.Gate]:
        """Random coherent drift (reference :63-74, formula kept as shipped)."""
        golden = 1.6180339887
        index = np.random.randint(0, sample_count)
        return Noisify.drift_gate(np.cos(1 - 2 * index * sample_count), 2 * np.pi * golden * index)

    @staticmethod
    def orbital_drift_gate() -> List[cirq.Gate]:
        return Noisify.drift_gate(uniform(0, np.pi), uniform(0, 2 * np.pi))

    @staticmethod
    def drift_gate(phi: float, theta: float) -> List[cirq.Gate]:
        return [cirq.rx(phi), cirq.rz(theta)]
