"""Compute halves of the reference's experiment drivers (``src/visual_testing.py``), without the plotting.

``visual_testing.py`` is a matplotlib/qiskit script and stays out of scope, but the
``overwrite=True`` branches of its experiments are the *callers* that produced the reference's stored results
(``src/classic_minimisation/{SWAP,H2}_Similarity_*``): they define the batch axes of the hot path and are the only
reference-generated outputs a noisy / mitigated value can be compared with.  Each function here returns the same
container object the reference writes with ``write_object`` (so ``calculate_minimisation.write_object`` stores
it in the same jsonpickle format).
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import cirq_compat as cirq
from .data_containers.helper_interfaces.i_collection import ISimilarityCollection
from .data_containers.helper_interfaces.i_noise_wrapper import INoiseModel, INoiseWrapper
from .data_containers.model_hydrogen import HydrogenAnsatz
from .error_mitigation import IErrorMitigationManager, SingleQubitPauliChannel, TwoQubitPauliChannel
from .processors.processor_classic import CPU
from .processors.processor_quantum import QPU
from .qp_qem_lib import swap_circuit


def asymmetric_pauli_noise_model(prob: float) -> INoiseModel:
    """The noise model of every stored experiment: p_x = p_y = prob, p_z = 6 prob on one- and two-qubit gates
    (visual_testing.py:496-498, 567-569)."""
    channel_1q = [SingleQubitPauliChannel(p_x=prob, p_y=prob, p_z=6 * prob)]
    channel_2q = [TwoQubitPauliChannel(p_x=prob, p_y=prob, p_z=6 * prob)]
    return INoiseModel(noise_gates_1q=channel_1q, noise_gates_2q=channel_2q, description=f'asymmetric depolarization (p_tot={16 * prob})')


def similarity_swap_circuit(prob: float = 1e-4, measure_count: int = 100, identifier_count: int = 1000,
                            density_representation: bool = False) -> ISimilarityCollection:
    """``similarity_plot_swap_circuit(overwrite=True)`` (visual_testing.py:492-524): 3+3-qubit SWAP test,
    ``measure_count`` repetitions of (noisy, QP-mitigated) estimates from ``identifier_count`` sampled circuits."""
    noise_model = asymmetric_pauli_noise_model(prob)
    n = 3
    circuit = swap_circuit(n, cirq.LineQubit.range(2 * n + 1), True)
    manager = IErrorMitigationManager(clean_circuit=circuit, noise_model=INoiseModel.empty(), hamiltonian_objective=None)
    manager.set_identifier_range(identifier_count)
    mu_ideal = manager.get_mu_effective(error_mitigation=False, density_representation=True, meas_reps=1)[1]
    manager = IErrorMitigationManager(clean_circuit=circuit, noise_model=noise_model, hamiltonian_objective=None)
    result_noisy, result_mitigated = [], []
    for _ in range(measure_count):
        manager.set_identifier_range(identifier_count)
        result_noisy.append(manager.get_mu_effective(error_mitigation=False, density_representation=density_representation, meas_reps=1)[1])
        result_mitigated.append(manager.get_mu_effective(error_mitigation=True, density_representation=density_representation, meas_reps=1)[1])
    settings = ISimilarityCollection.get_settings(probability=prob, measure_count=measure_count, identifier_count=identifier_count)
    return ISimilarityCollection(data_noisy=result_noisy, data_mitigated=result_mitigated, mu_ideal=mu_ideal, settings=settings)


def similarity_hydrogen_circuit(prob: float = 1e-4, measure_count: int = 100, identifier_count: int = 1000,
                                max_cpu_iter: int = 10, max_qpu_iter: int = 10000,
                                fixed_parameters: Optional[np.ndarray] = None) -> ISimilarityCollection:
    """``similarity_plot_hydrogen_circuit(overwrite=True)`` (visual_testing.py:562-607): per repetition optimise the
    noisy ansatz with shot-based COBYLA, then estimate the noisy and the QP-mitigated energy with the sampled
    ``measure_observable`` objective.  ``fixed_parameters`` skips the optimisation (not in the reference)."""
    noise_model = asymmetric_pauli_noise_model(prob)
    ansatz = HydrogenAnsatz()
    noisy_parameters = ansatz.operator_parameters
    noisy_ansatz = INoiseWrapper(w_class=ansatz, noise_model=noise_model)
    circuit = noisy_ansatz.get_clean_circuit()
    mu_ideal = -1.13727  # hard-coded in the reference (visual_testing.py:590)
    result_noisy, result_mitigated = [], []
    for _ in range(measure_count):
        if fixed_parameters is None:
            _exp_value, opt_params = CPU.get_custom_optimized_state(n_w=noisy_ansatz, max_cpu_iter=max_cpu_iter, max_qpu_iter=max_qpu_iter)
        else:
            opt_params = np.asarray(fixed_parameters, dtype=float)
        noisy_parameters.update(v=opt_params)
        circuit_noisy_param = QPU.get_resolved_circuit(c=circuit, p=noisy_parameters)
        manager = IErrorMitigationManager(clean_circuit=circuit_noisy_param, noise_model=noise_model, hamiltonian_objective=ansatz)
        manager.set_identifier_range(identifier_count)
        result_noisy.append(manager.get_mu_effective(error_mitigation=False, density_representation=False, meas_reps=1)[1])
        result_mitigated.append(manager.get_mu_effective(error_mitigation=True, density_representation=False, meas_reps=1)[1])
    settings = ISimilarityCollection.get_settings(probability=prob, measure_count=measure_count, identifier_count=identifier_count)
    return ISimilarityCollection(data_noisy=result_noisy, data_mitigated=result_mitigated, mu_ideal=mu_ideal, settings=settings)
