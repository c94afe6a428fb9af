"""Demo drivers (mirror of ``src/main.py``): ``custom_ansatz`` (:35-50), ``basic_ansatz`` (:12-32, broken in the
reference as shipped because ``uccsd_ansatz.circuit`` no longer exists; here it uses ``operations(qubits)``),
``get_log_experiment`` (:53-62)."""
from __future__ import annotations

from typing import Callable, Dict, List

import numpy as np

from . import cirq_compat as cirq
from .data_containers.helper_interfaces.i_noise_wrapper import INoiseModel, INoiseWrapper
from .data_containers.model_hydrogen import HydrogenAnsatz
from .processors.processor_classic import CPU
from .processors.processor_quantum import QPU


def basic_ansatz(max_iter: int = 1000):
    uccsd_ansatz = HydrogenAnsatz()
    parameters = uccsd_ansatz.operator_parameters
    circuit = QPU.get_initial_state_circuit(uccsd_ansatz)
    circuit.append(uccsd_ansatz.operations(uccsd_ansatz.qubits))
    print(circuit)
    result = CPU.get_optimized_state(w=uccsd_ansatz, max_iter=max_iter)
    print(f'Operator expectation value: {result.optimal_value}\nOperator parameters: {result.optimal_parameters}')
    parameters.update(v=result.optimal_parameters)
    print(QPU.get_resolved_circuit(circuit, parameters))
    return result


def custom_ansatz(max_cpu_iter: int = 10, max_qpu_iter: int = 10000, p: float = .5):
    uccsd_ansatz = HydrogenAnsatz()
    parameters = uccsd_ansatz.operator_parameters
    noise_channel = INoiseModel(noise_gates_1q=[cirq.bit_flip(p=p), cirq.phase_flip(p=p)], noise_gates_2q=[],
                                description=f'Bit and Phase flip (p={p})')
    noisy_ansatz = INoiseWrapper(uccsd_ansatz, noise_channel)
    values, params = CPU.get_custom_optimized_state(n_w=noisy_ansatz, max_cpu_iter=max_cpu_iter, max_qpu_iter=max_qpu_iter)
    print(f'Operator expectation value: {values}\nOperator parameters: {params}')
    for i, key in enumerate(parameters.dict.keys()):
        parameters[key] = params[i]
    print(QPU.get_resolved_circuit(noisy_ansatz.get_noisy_circuit(), parameters))
    return values, params


def get_log_experiment(shot_list: List[int], expectation: float, experiment: Callable[[int], float], reps: int = 1) -> Dict[int, float]:
    """|experiment(shots) - expectation| averaged over ``reps`` repetitions, per entry of ``shot_list``."""
    result_dict = {}
    for shots in shot_list:
        result_dict[shots] = np.average([np.abs(experiment(shots) - expectation) for _ in range(reps)])
    return result_dict


if __name__ == '__main__':
    custom_ansatz()
