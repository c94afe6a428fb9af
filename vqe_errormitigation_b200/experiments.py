"""Compute halves of the reference's experiment drivers (``src/visual_testing.py``), without the plotting.

``visual_testing.py`` is a matplotlib/qiskit script and stays out of scope, but the
``overwrite=True`` branches of its experiments are the *callers* that produced the reference's stored results
(``src/classic_minimisation/{SWAP,H2}_Similarity_*``): they define the batch axes of the hot path and are the only
reference-generated outputs a noisy / mitigated value can be compared with.  Each function here returns the same
container object the reference writes with ``write_object`` (so ``calculate_minimisation.write_object`` stores
it in the same jsonpickle format).
"""
from __future__ import annotations

from typing import Optional, Sequence

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


def simulate_density_matrix(circuit: cirq.Circuit) -> np.ndarray:
    """Final density matrix with measurements treated as dephasing (visual_testing.py:109-112)."""
    return cirq.DensityMatrixSimulator(ignore_measurement_results=True).simulate(program=circuit).final_density_matrix


def sampling_noise_scaling(shot_list: Optional[Sequence[int]] = None, reps: int = 10, prob: float = 1e-4, max_iter: int = 1000):
    """``sampling_noise_scaling`` (visual_testing.py:334-387): error of the QP-mitigated Tr(rho H) of the optimised H2 ansatz
    against the number of sampled circuits.  Returns ({count: mean |estimate - optimum| over ``reps``}, (slope, intercept) of
    the reference's linear fit of the error over log10(count), or None when the fit fails).  Every estimate is one
    variant batch; the reference omits ``meas_reps`` in its call (a TypeError as shipped) - 1 here, unused on the density path."""
    from .main import get_log_experiment
    ansatz = HydrogenAnsatz()
    noise_model = asymmetric_pauli_noise_model(prob)
    circuit = INoiseWrapper(ansatz, noise_model).get_clean_circuit()
    result = CPU.get_optimized_state(w=ansatz, max_iter=max_iter)
    parameters = ansatz.operator_parameters
    parameters.update(v=result.optimal_parameters)
    resolved_circuit = QPU.get_resolved_circuit(circuit, parameters)
    observable = QPU.get_hamiltonian_objective_operator(ansatz)      # scipy sparse: the manager takes Tr(rho H) with the full H

    def experiment(process_circuit_count: int) -> float:
        return CPU.get_mitigated_expectation(clean_circuit=resolved_circuit, noise_model=noise_model, process_circuit_count=process_circuit_count,
                                             hamiltonian_objective=observable, meas_reps=1)

    shot_list = [int(shot) for shot in np.logspace(1, 4, 10)] if shot_list is None else [int(shot) for shot in shot_list]
    errors = get_log_experiment(shot_list=shot_list, expectation=result.optimal_value, experiment=experiment, reps=reps)
    try:
        fit = tuple(float(c) for c in np.polyfit(np.log10(np.array(shot_list)), np.array(list(errors.values())), 1))
    except (np.linalg.LinAlgError, TypeError):
        fit = None
    return errors, fit


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


def hamiltonian_density_vs_measure(prob_list: Optional[Sequence[float]] = None, meas_reps: Optional[Sequence[int]] = None, count: int = 50,
                                   max_iter: int = 1000):
    """``hamiltonian_density_vs_measure`` (visual_testing.py:402-447): the deterministic Tr(rho H) of the noisy optimised
    ansatz against ``count`` shot-based ``measure_observable`` energies per (noise level, shots).  Returns
    (error[i, j] = |Tr(rho H) - mean|, spread[i, j] = std, prob_list, meas_reps); the reference's Pool(6) over the
    ``count`` repetitions (:438-439) becomes one batch per (noise level, shots): ``count`` parameter rows x five measurement circuits
    in one launch, shots drawn on the device (``SampledEnergyProgram.energies``)."""
    ansatz = HydrogenAnsatz()
    result = CPU.get_optimized_state(w=ansatz, max_iter=max_iter)
    parameters = ansatz.operator_parameters
    parameters.update(v=result.optimal_parameters)       # (the reference passes r=result, which its IParameter rejects)
    observable_process, _circuit_cost = ansatz.observable_measurement()
    prob_list = [float(1 / shot) for shot in np.logspace(1, 4, 4)] if prob_list is None else list(prob_list)
    meas_reps = [int(shot) for shot in np.logspace(0, 4, 10)] if meas_reps is None else list(meas_reps)
    error_list = np.zeros((len(prob_list), len(meas_reps)))
    error_var_list = np.zeros((len(prob_list), len(meas_reps)))
    for i, prob in enumerate(prob_list):
        noise_model = asymmetric_pauli_noise_model(prob)
        noisy_ansatz = INoiseWrapper(ansatz, noise_model)
        circuit = noisy_ansatz.get_noisy_circuit()
        exp_fidelity = QPU.get_simulated_noisy_expectation_value(w=noisy_ansatz, r_c=circuit, r=parameters.get_resolved())
        program = noisy_ansatz.sampled_energy_program(circuit)
        names = list(parameters)
        row = np.array([[parameters[nm] for nm in program.prog.param_names]], dtype=float) if program.prog.param_names else None
        assert row is None or set(program.prog.param_names) <= set(names)
        for j, meas in enumerate(meas_reps):
            # the ``count`` repetitions are rows of one batch (own Philox streams), not ``count`` calls of observable_process
            seed = int(np.random.randint(0, 2 ** 31 - 1))
            if row is None:
                exp_measure = np.array([observable_process(circuit, noise_model.get_operators, meas) for _ in range(count)])
            else:
                exp_measure = program.energies(np.repeat(row, count, axis=0), meas, seed)
            error_list[i][j] = abs(exp_fidelity - np.mean(exp_measure))
            error_var_list[i][j] = abs(np.std(exp_measure))
    return error_list, error_var_list, prob_list, meas_reps


def circuit_parameter_optimization(prob_list: Optional[Sequence[float]] = None, count: int = 20, max_cpu_iter: int = 10,
                                   max_qpu_iter: int = 10000):
    """``circuit_parameter_optimization(overwrite=True)`` (visual_testing.py:660-700): ``count`` independent shot-based
    COBYLA runs per noise level.  Returns the tuple the reference stores as ``H2_Optimizing_Parameters_02``:
    (exp_list, opt_list, prob_list)."""
    ansatz = HydrogenAnsatz()
    prob_list = np.flip(np.array([1 / shot for shot in np.logspace(2, 5, 14)])) if prob_list is None else np.asarray(prob_list, float)
    exp_list, opt_list = [], []
    for prob in prob_list:
        noisy_ansatz = INoiseWrapper(w_class=ansatz, noise_model=asymmetric_pauli_noise_model(float(prob)))
        # the reference maps the runs over Pool(6) workers, each with its own pickled copy of the wrapper: ``count`` independent
        # COBYLA runs from the same start point - here lock-step chains, one launch per optimiser round
        values, params = CPU.get_custom_optimized_states_lockstep(noisy_ansatz, restarts=count, max_cpu_iter=max_cpu_iter, max_qpu_iter=max_qpu_iter)
        exp_list.append([float(v) for v in values])
        opt_list.append([float(x[0]) for x in params])
    return exp_list, opt_list, prob_list


def raw_and_mitigated(n_w: INoiseWrapper, noise_model: INoiseModel, max_qpu_iter: int, max_cpu_iter: int = 10):
    """``RawAndMitigatedClass.process`` (visual_testing.py:746-771): shot-based noisy optimum, then the QP-mitigated energy
    at the optimised parameters (objective = the noise wrapper itself, sampled ``measure_observable``)."""
    exp_noisy_value, opt_params = CPU.get_custom_optimized_state(n_w=n_w, max_cpu_iter=max_cpu_iter, max_qpu_iter=max_qpu_iter)
    ansatz_parameters = n_w.operator_parameters
    ansatz_parameters.update(v=opt_params)
    circuit_noisy_param = QPU.get_resolved_circuit(c=n_w.get_clean_circuit(), p=ansatz_parameters)
    manager = IErrorMitigationManager(clean_circuit=circuit_noisy_param, noise_model=noise_model, hamiltonian_objective=n_w)
    manager.set_identifier_range(max_qpu_iter)
    exp_mitig_value = manager.get_mu_effective(error_mitigation=True, density_representation=False, meas_reps=1)[1]
    return exp_noisy_value, exp_mitig_value


def raw_and_mitigated_repeated(n_w: INoiseWrapper, noise_model: INoiseModel, repetitions: int, max_qpu_iter: int, max_cpu_iter: int = 10):
    """``repetitions`` x :func:`raw_and_mitigated` with the noisy optimisations advanced in lock-step (one launch per COBYLA
    round); each optimum then gets its QP-mitigated estimate.  NOT the reference's order of events: its drivers call
    ``process()`` in a serial loop on one shared wrapper (visual_testing.py:800-801,893-894), so every repetition starts from
    the previous repetition's optimum; here every repetition starts from the parameters the wrapper holds on entry, which are
    restored on exit.  Opt-in through ``lockstep=True`` of the two drivers below.  Returns [(noisy, mitigated)] * repetitions."""
    ansatz_parameters = n_w.operator_parameters
    start = dict(ansatz_parameters.dict)
    values, params = CPU.get_custom_optimized_states_lockstep(n_w, restarts=repetitions, max_cpu_iter=max_cpu_iter, max_qpu_iter=max_qpu_iter)
    results = []
    try:
        for exp_noisy_value, opt_params in zip(values, params):
            ansatz_parameters.update(v=opt_params)
            circuit_noisy_param = QPU.get_resolved_circuit(c=n_w.get_clean_circuit(), p=ansatz_parameters)
            manager = IErrorMitigationManager(clean_circuit=circuit_noisy_param, noise_model=noise_model, hamiltonian_objective=n_w)
            manager.set_identifier_range(max_qpu_iter)
            results.append((float(exp_noisy_value), manager.get_mu_effective(error_mitigation=True, density_representation=False, meas_reps=1)[1]))
    finally:
        ansatz_parameters.dict.update(start)
    return results


def full_raw_vs_mitigation_per_noise(measure_count: int, identifier_count: int, prob_list: Optional[Sequence[float]] = None,
                                     max_cpu_iter: int = 10, lockstep: bool = False):
    """``full_raw_vs_mitigation_per_noise(overwrite=True)`` (visual_testing.py:778-812): per noise level ``measure_count``
    x (noisy, mitigated).  Returns (exp_noisy_list, exp_mitig_list, prob_list) as stored in ``H2_Measure_Raw_vs_Mitigated``."""
    ansatz = HydrogenAnsatz()
    prob_list = np.flip(np.array([1 / shot for shot in np.logspace(3, 5, 5)])) if prob_list is None else np.asarray(prob_list, float)
    exp_noisy_list, exp_mitig_list = [], []
    for prob in prob_list:
        noise_model = asymmetric_pauli_noise_model(float(prob))
        noisy_ansatz = INoiseWrapper(w_class=ansatz, noise_model=noise_model)
        results = raw_and_mitigated_repeated(noisy_ansatz, noise_model, measure_count, identifier_count, max_cpu_iter) if lockstep else \
            [raw_and_mitigated(noisy_ansatz, noise_model, identifier_count, max_cpu_iter) for _ in range(measure_count)]
        exp_noisy_list.append([noisy for noisy, _ in results])
        exp_mitig_list.append([mitigated for _, mitigated in results])
    return exp_noisy_list, exp_mitig_list, prob_list


def hydrogen_energy_spectrum(measure_count: int, identifier_count: int, atomic_distance_list: Optional[Sequence[float]] = None,
                             previous=None, prob: float = 1e-4, max_cpu_iter: int = 10, lockstep: bool = False):
    """``hydrogen_energy_spectrum(overwrite=True)`` (visual_testing.py:863-909): (noisy, mitigated) energies along the bond
    length, ``measure_count`` repetitions per point; ``previous`` is an earlier result tuple to extend (the reference
    re-reads ``H2_Full_Spectrum_C<n>`` and appends, :866-871).  Returns (exp_noisy_list, exp_mitig_list,
    atomic_distance_list, fci_energy, hf_energy)."""
    atomic_distance_list = np.linspace(.1, 2.8, 10) if atomic_distance_list is None else np.asarray(atomic_distance_list, float)
    depth = len(atomic_distance_list)
    exp_noisy_list = [list(v) for v in previous[0]] if previous else [[] for _ in range(depth)]
    exp_mitig_list = [list(v) for v in previous[1]] if previous else [[] for _ in range(depth)]
    noise_model = asymmetric_pauli_noise_model(prob)
    fci_energy, hf_energy = [], []
    for i, atomic_distance in enumerate(atomic_distance_list):
        ansatz = HydrogenAnsatz(mol_param=round(float(atomic_distance), 2))
        noisy_ansatz = INoiseWrapper(w_class=ansatz, noise_model=noise_model)
        results = raw_and_mitigated_repeated(noisy_ansatz, noise_model, measure_count, identifier_count, max_cpu_iter) if lockstep else \
            [raw_and_mitigated(noisy_ansatz, noise_model, identifier_count, max_cpu_iter) for _ in range(measure_count)]
        exp_noisy_list[i].extend(noisy for noisy, _ in results)
        exp_mitig_list[i].extend(mitigated for _, mitigated in results)
        fci_energy.append(float(ansatz.molecule.fci_energy))
        hf_energy.append(float(ansatz.molecule.hf_energy))
    return exp_noisy_list, exp_mitig_list, atomic_distance_list, fci_energy, hf_energy


def hydrogen_fci_energy(atomic_distance_list: Optional[Sequence[float]] = None, repetitions: int = 10, max_cpu_iter: int = 10,
                        max_qpu_iter: int = 1000):
    """``hydrogen_fci_energy(overwrite=True)`` (visual_testing.py:958-990): noiseless shot-based VQE energies along the bond
    length.  Bond lengths without (complete) molecular data give NaN rows, like the reference's bare ``except`` (:983-986):
    r = 0.0 has no hdf5 at all, ``H2_1.8.hdf5`` stops after the SCF step.  Returns (atomic_distance_list, fci_energy,
    hf_energy, measured_energy)."""
    atomic_distance_list = np.linspace(0.0, 3.0, 31) if atomic_distance_list is None else np.asarray(atomic_distance_list, float)
    noise_model = asymmetric_pauli_noise_model(0)
    fci_energy, hf_energy, measured_energy = [], [], []
    for atomic_distance in atomic_distance_list:
        try:
            ansatz = HydrogenAnsatz(mol_param=round(float(atomic_distance), 2))
            noisy_ansatz = INoiseWrapper(ansatz, noise_model)
            energies = [CPU.get_custom_optimized_state(n_w=noisy_ansatz, max_cpu_iter=max_cpu_iter, max_qpu_iter=max_qpu_iter)[0]
                        for _ in range(repetitions)]
            fci, hf = float(ansatz.molecule.fci_energy), float(ansatz.molecule.hf_energy)
        except (FileNotFoundError, TypeError, ValueError, KeyError):
            energies, fci, hf = [float('NaN')] * repetitions, float('NaN'), float('NaN')
        measured_energy.append(energies)
        fci_energy.append(fci)
        hf_energy.append(hf)
    return atomic_distance_list, fci_energy, hf_energy, measured_energy
