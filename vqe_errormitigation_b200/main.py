"""Demo drivers with the names of ``src/main.py``: ``custom_ansatz`` (:35-50: shot-based noisy optimisation of the H2
ansatz under a bit + phase flip model), ``basic_ansatz`` (:12-32: noiseless optimisation; broken in the reference as
shipped because ``uccsd_ansatz.circuit`` no longer exists — here the circuit comes from ``operations(qubits)``) and
``get_log_experiment`` (:53-62)."""
from __future__ import annotations

from typing import Callable, Dict, List

import numpy as np

from . import cirq_compat as cirq
from .data_containers.helper_interfaces.i_noise_wrapper import INoiseModel, INoiseWrapper
from .data_containers.model_hydrogen import HydrogenAnsatz
from .processors.processor_classic import CPU
from .processors.processor_quantum import QPU


def _report(value, params, circuit):
    print(f"expectation value {value}\nparameters {params}")
    print(circuit)


def basic_ansatz(max_iter: int = 1000):
    ansatz = HydrogenAnsatz()
    circuit = QPU.get_initial_state_circuit(ansatz)
    circuit.append(ansatz.operations(ansatz.qubits))
    result = CPU.get_optimized_state(w=ansatz, max_iter=max_iter)
    ansatz.operator_parameters.update(v=result.optimal_parameters)
    _report(result.optimal_value, result.optimal_parameters, QPU.get_resolved_circuit(circuit, ansatz.operator_parameters))
    return result


def custom_ansatz(max_cpu_iter: int = 10, max_qpu_iter: int = 10000, p: float = .5):
    ansatz = HydrogenAnsatz()
    flips = INoiseModel(noise_gates_1q=[cirq.bit_flip(p=p), cirq.phase_flip(p=p)], noise_gates_2q=[],
                        description=f"Bit and Phase flip (p={p})")
    noisy = INoiseWrapper(ansatz, flips)
    value, params = CPU.get_custom_optimized_state(n_w=noisy, max_cpu_iter=max_cpu_iter, max_qpu_iter=max_qpu_iter)
    table = ansatz.operator_parameters               # shared with the wrapper: write the optimum back by position
    for name, optimum in zip(list(table.dict), params):
        table[name] = optimum
    _report(value, params, QPU.get_resolved_circuit(noisy.get_noisy_circuit(), table))
    return value, params


def get_log_experiment(shot_list: List[int], expectation: float, experiment: Callable[[int], float], reps: int = 1) -> Dict[int, float]:
    """shots -> mean |experiment(shots) - expectation| over ``reps`` repetitions."""
    return {shots: float(np.mean([abs(experiment(shots) - expectation) for _ in range(reps)])) for shots in shot_list}


if __name__ == "__main__":
    custom_ansatz()
