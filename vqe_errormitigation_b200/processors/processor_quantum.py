"""``QPU``: the "quantum processor" facade (mirror of ``src/processors/processor_quantum.py``).

Same static methods and argument meaning as the reference; the density-matrix evolution and Tr(rho H) run
on the GPU.  Differences, on purpose: FP64 throughout (the reference's cirq default is complex64), and the
Jordan-Wigner Hamiltonian is cached per molecule instead of being re-read from disk on every call
(reference :74-78 reloads the hdf5 and redoes JW per evaluation).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np

from .. import cirq_compat as cirq
from .. import engine
from ..data_containers.helper_interfaces.i_parameter import IParameter
from ..data_containers.helper_interfaces.i_wave_function import IWaveFunction
from ..openfermion_compat import QubitOperator, get_sparse_operator, jordan_wigner, pauli_masks


class QPU:
    _hamiltonian_cache: Dict[Tuple, QubitOperator] = {}

    @staticmethod
    def get_hamiltonian_evaluation_operator(w: IWaveFunction) -> QubitOperator:
        molecule = w.molecule
        key = (id(molecule), molecule.description, molecule.nuclear_repulsion)
        cached = QPU._hamiltonian_cache.get(key)
        if cached is None:
            if molecule.one_body_integrals is None:
                molecule.load()
            cached = jordan_wigner(molecule.get_molecular_hamiltonian())
            if len(QPU._hamiltonian_cache) > 256:
                QPU._hamiltonian_cache.clear()
            QPU._hamiltonian_cache[key] = cached
        return cached

    @staticmethod
    def get_hamiltonian_objective_operator(w: IWaveFunction):
        return get_sparse_operator(QPU.get_hamiltonian_evaluation_operator(w), n_qubits=len(w.qubits))

    @staticmethod
    def get_hamiltonian_terms(w: IWaveFunction):
        """(xmask, zmask, coeff) arrays of the Hamiltonian over ``w.qubits`` (engine input format)."""
        return pauli_masks(QPU.get_hamiltonian_evaluation_operator(w), len(w.qubits))

    @staticmethod
    def get_simulated_noisy_expectation_value(w: IWaveFunction, r_c: cirq.Circuit, r) -> float:
        """Re Tr(rho H) of the (resolved) circuit — reference :28-39."""
        resolved = cirq.resolve_parameters(r_c, r) if r is not None else r_c
        prog, params = engine.program_for_circuit(resolved, w.qubits, hamiltonian=QPU.get_hamiltonian_terms(w))
        return float(prog.energies(params, batch=1).cpu().numpy()[0])

    @staticmethod
    def get_realistic_noisy_expectation_value(n_w, c: cirq.Circuit, p: IParameter, max_iter: int) -> float:
        circuit_to_run = QPU.get_resolved_circuit(c=c, p=p)
        _func, _cost = n_w.observable_measurement()
        return _func(circuit_to_run, n_w._noise_channel, max_iter)

    @staticmethod
    def get_expectation_value(t_r, w: IWaveFunction):
        raise NotImplementedError("dead in the reference as shipped (objective.value on a sparse matrix, processor_quantum.py:16-19)")

    @staticmethod
    def get_trial_results(r_c: cirq.Circuit, r, max_iter: int):
        """Reference :48-52: ``cirq.Simulator().run`` (the facade samples the noiseless density-matrix evolution: same
        distribution)."""
        return cirq.Simulator().run(program=r_c, param_resolver=r, repetitions=max_iter)

    @staticmethod
    def get_noisy_trial_result(r_c: cirq.Circuit, r, max_iter: int, noise_model=None):
        """Reference :55-56: ``cirq.sample(program, noise, param_resolver, repetitions)``."""
        return cirq.sample(program=r_c, noise=noise_model, param_resolver=r, repetitions=max_iter)

    @staticmethod
    def get_initial_state_circuit(w: IWaveFunction) -> cirq.Circuit:
        return cirq.Circuit(w.initial_state(qubits=w.qubits), strategy=cirq.InsertStrategy.EARLIEST)

    @staticmethod
    def get_resolved_circuit(c: cirq.Circuit, p: IParameter) -> cirq.Circuit:
        return cirq.resolve_parameters(c, p.get_resolved())

    # ---- batched entry points (not in the reference) ---------------------------------------
    @staticmethod
    def compile_energy_program(w: IWaveFunction, circuit: cirq.Circuit, **options) -> engine.CircuitProgram:
        """Compile a (symbolic) circuit + the molecule's Hamiltonian once; evaluate many parameter sets with
        ``program.energies(params[B, P])`` (columns ordered as ``program.param_names``)."""
        prog, _, _ = engine.compile_circuit(circuit, w.qubits, hamiltonian=QPU.get_hamiltonian_terms(w), options=options)
        return prog

    @staticmethod
    def get_simulated_noisy_expectation_values(w: IWaveFunction, circuit: cirq.Circuit, params: np.ndarray) -> np.ndarray:
        """Batched form of get_simulated_noisy_expectation_value: ``params[B, P]`` in the order of
        ``w.operator_parameters``; returns energies[B] (numpy)."""
        prog = QPU.compile_energy_program(w, circuit)
        names = list(w.operator_parameters)
        params = np.asarray(params, dtype=np.float64).reshape(-1, len(names))
        cols = [names.index(nm) for nm in prog.param_names]
        return prog.energies_host(params[:, cols] if cols else None, batch=params.shape[0])
