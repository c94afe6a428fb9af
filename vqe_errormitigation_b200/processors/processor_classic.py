"""``CPU``: classical optimiser + error-mitigation orchestration (mirror of ``src/processors/processor_classic.py``).

The single-chain COBYLA path keeps the reference's behaviour (start from the current parameter values,
``tol=1e-6``, last-evaluated values left in the parameter dict).  ``get_batched_optimized_states`` is the
B200 addition: many independent COBYLA chains advance in lock-step and every round of objective requests is
ONE batched energy launch (scipy's COBYLA is strictly one-evaluation-at-a-time).
"""
from __future__ import annotations

import copy
import threading
from typing import Callable, List, Optional, Sequence, Tuple, Union

import numpy as np
from scipy import optimize

from .. import cirq_compat as cirq
from ..data_containers.helper_interfaces.i_collection import IContainer, IMeasurementCollection
from ..data_containers.helper_interfaces.i_noise_wrapper import INoiseModel, INoiseWrapper
from ..data_containers.helper_interfaces.i_parameter import IParameter
from ..data_containers.helper_interfaces.i_wave_function import IWaveFunction
from ..error_mitigation import IErrorMitigationManager
from .processor_quantum import QPU


class OptimizationTrialResult:
    """Minimal stand-in for openfermioncirq's result object (``.optimal_value`` / ``.optimal_parameters``)."""

    def __init__(self, optimal_value: float, optimal_parameters: np.ndarray, num_evaluations: int = 0):
        self.optimal_value = optimal_value
        self.optimal_parameters = optimal_parameters
        self.num_evaluations = num_evaluations


class CPU:
    @staticmethod
    def get_collection_ground_states(collection: IMeasurementCollection, cpu_iter: int, qpu_iter: int) -> List[IContainer]:
        """noise models x bond lengths x ``cpu_iter`` restarts (reference :18-44)."""
        result = []
        for wave_function, description in collection.get_wave_functions():
            values = []
            for _ in range(cpu_iter):
                if isinstance(wave_function, INoiseWrapper):
                    value, par = CPU.get_custom_optimized_state(n_w=copy.deepcopy(wave_function), max_cpu_iter=cpu_iter, max_qpu_iter=qpu_iter)
                    print(f"par: {par}, val: {value}")
                else:
                    value = CPU.get_optimized_state(w=wave_function, max_iter=qpu_iter).optimal_value
                values.append(value)
            result.append(IContainer(m_param=wave_function.molecule_parameters, e_values=values, m_data=wave_function.molecule, label=description))
        return result

    @staticmethod
    def get_optimized_state(w: IWaveFunction, max_iter: int) -> OptimizationTrialResult:
        """Noiseless optimisation of <H> over the ansatz parameters.  The reference's version (openfermioncirq
        VariationalStudy, processor_classic.py:46-50) is commented out yet still called; this restores it with
        the deterministic Tr(rho H) objective and COBYLA."""
        circuit = QPU.get_initial_state_circuit(w=w)
        circuit.append(w.operations(w.qubits))
        prog = QPU.compile_energy_program(w, circuit)
        names = list(w.operator_parameters)
        cols = [names.index(nm) for nm in prog.param_names]

        def objective(x: np.ndarray) -> float:
            return float(prog.energies_host(np.asarray(x, float)[None, cols] if cols else None, batch=1)[0])

        x0 = np.fromiter(w.operator_parameters.dict.values(), dtype=float)
        res = optimize.minimize(fun=objective, x0=x0, method="COBYLA", options={"maxiter": max_iter})
        return OptimizationTrialResult(float(res.fun), np.atleast_1d(res.x), int(getattr(res, "nfev", 0)))

    @staticmethod
    def get_custom_optimized_state(n_w: INoiseWrapper, max_cpu_iter: int, max_qpu_iter: int, use_density: bool = False) -> Tuple[float, np.ndarray]:
        """COBYLA over the noisy circuit's parameters (reference :52-92).  Objective = the sampled-energy
        function of ``observable_measurement`` with ``max_qpu_iter`` shots per circuit; ``use_density=True``
        switches to the deterministic Tr(rho H) objective the reference keeps commented out at :73-75."""
        operator_params = n_w.operator_parameters
        evaluation_circuit = n_w.get_noisy_circuit()
        expectation_function, _cost = n_w.observable_measurement()

        def optimize_func(x: np.ndarray) -> float:
            for i, key in enumerate(operator_params):
                operator_params.dict[key] = x[i]
            if use_density:
                return QPU.get_simulated_noisy_expectation_value(w=n_w, r_c=evaluation_circuit, r=operator_params.get_resolved())
            circuit_to_run = QPU.get_resolved_circuit(c=evaluation_circuit, p=operator_params)
            return expectation_function(circuit_to_run, n_w._noise_channel, max_qpu_iter)

        x0 = np.fromiter(operator_params.dict.values(), dtype=float)
        res = optimize.minimize(fun=optimize_func, x0=x0, method="COBYLA", options={"maxiter": max_cpu_iter}, tol=1e-6)
        return res.fun, res.x

    @staticmethod
    def get_mitigated_optimized_state(w: IWaveFunction, n: INoiseModel, max_optimization_iter: int, max_mitigation_count: int) -> float:
        """Optimise the clean ansatz, then estimate the mitigated energy (reference :94-125)."""
        circuit_parameters = w.operator_parameters
        noisy_ansatz = INoiseWrapper(w_class=w, noise_model=n)
        result = CPU.get_optimized_state(w=w, max_iter=max_optimization_iter)
        circuit_parameters.update(v=result.optimal_parameters)
        resolved_circuit = QPU.get_resolved_circuit(noisy_ansatz.get_clean_circuit(), circuit_parameters)
        return CPU.get_mitigated_expectation(clean_circuit=resolved_circuit, noise_model=n, process_circuit_count=max_mitigation_count,
                                             hamiltonian_objective=w, meas_reps=1)

    @staticmethod
    def get_mitigated_expectation(clean_circuit: cirq.Circuit, noise_model: INoiseModel, process_circuit_count: int,
                                  hamiltonian_objective: Union[np.ndarray, IWaveFunction, None], meas_reps: int) -> float:
        manager = IErrorMitigationManager(clean_circuit=clean_circuit, noise_model=noise_model, hamiltonian_objective=hamiltonian_objective)
        manager.set_identifier_range(process_circuit_count)
        using_density = not isinstance(hamiltonian_objective, IWaveFunction)
        _values, mu = manager.get_mu_effective(error_mitigation=True, density_representation=using_density, meas_reps=meas_reps)
        return mu

    # ---- batched additions ----------------------------------------------------------------------------------------
    @staticmethod
    def get_batched_energies(n_w: IWaveFunction, params: np.ndarray, noisy: bool = True) -> np.ndarray:
        """Tr(rho H) for ``params[B, P]`` (columns in ``operator_parameters`` order) in one launch."""
        circuit = n_w.get_noisy_circuit() if (noisy and isinstance(n_w, INoiseWrapper)) else None
        if circuit is None:
            circuit = QPU.get_initial_state_circuit(w=n_w)
            circuit.append(n_w.operations(n_w.qubits))
        return QPU.get_simulated_noisy_expectation_values(n_w, circuit, params)

    @staticmethod
    def get_batched_optimized_states(n_w: IWaveFunction, x0s: np.ndarray, max_cpu_iter: int, tol: float = 1e-6,
                                     batch_objective: Optional[Callable[[np.ndarray], np.ndarray]] = None) -> Tuple[np.ndarray, np.ndarray]:
        """Run one COBYLA chain per row of ``x0s`` in lock-step; each round of objective requests is one batched
        Tr(rho H) launch.  Returns (fun[C], x[C, P]).  Each chain is scipy's own COBYLA (same trust-region
        schedule as the single-chain path); only the evaluation is batched."""
        x0s = np.atleast_2d(np.asarray(x0s, dtype=float))
        if batch_objective is None:
            circuit = n_w.get_noisy_circuit() if isinstance(n_w, INoiseWrapper) else None
            if circuit is None:
                circuit = QPU.get_initial_state_circuit(w=n_w)
                circuit.append(n_w.operations(n_w.qubits))
            prog = QPU.compile_energy_program(n_w, circuit)
            names = list(n_w.operator_parameters)
            cols = [names.index(nm) for nm in prog.param_names]
            batch_objective = lambda xs: prog.energies_host(np.ascontiguousarray(xs[:, cols]))
        return lockstep_minimize(batch_objective, x0s, max_cpu_iter, tol)


class _BatchFailed(Exception):
    """Raised inside a chain's objective after the coordinator's batched evaluation failed."""


def lockstep_minimize(batch_objective: Callable[[np.ndarray], np.ndarray], x0s: np.ndarray, maxiter: int, tol: float = 1e-6,
                      method: str = "COBYLA") -> Tuple[np.ndarray, np.ndarray]:
    """Advance ``len(x0s)`` independent scipy optimiser chains together: every chain runs in its own thread and
    blocks inside its objective until the coordinator has evaluated the whole round as one batch."""
    n_chains, dim = x0s.shape
    cond = threading.Condition()
    pending = {}          # chain -> requested x
    answers = {}          # chain -> value
    alive = set(range(n_chains))
    results: List[Optional[optimize.OptimizeResult]] = [None] * n_chains
    errors: List[BaseException] = []
    failed = threading.Event()

    def chain(c: int):
        def f(x):
            with cond:
                pending[c] = np.array(x, dtype=float, copy=True)
                cond.notify_all()
                while c not in answers:
                    cond.wait()
                value = answers.pop(c)
            if failed.is_set():
                raise _BatchFailed()      # unwinds scipy's optimiser; the coordinator re-raises the real error
            return value
        try:
            results[c] = optimize.minimize(fun=f, x0=x0s[c], method=method, options={"maxiter": maxiter}, tol=tol)
        except _BatchFailed:
            pass
        except BaseException as exc:  # noqa: BLE001 - propagated below
            errors.append(exc)
        finally:
            with cond:
                alive.discard(c)
                cond.notify_all()

    threads = [threading.Thread(target=chain, args=(c,), daemon=True) for c in range(n_chains)]
    for t in threads:
        t.start()
    while True:
        with cond:
            while alive and len(pending) < len(alive):
                cond.wait()
            if not alive:
                break
            ids = sorted(pending)
            xs = np.stack([pending.pop(c) for c in ids])
        try:
            vals = np.asarray(batch_objective(xs), dtype=float)
            if vals.shape != (len(ids),):
                raise ValueError(f"batch_objective returned shape {vals.shape} for {len(ids)} requests")
        except BaseException as exc:  # noqa: BLE001 - a failed batch must not leave the chains blocked in cond.wait()
            errors.insert(0, exc)
            vals = np.full(len(ids), np.nan)
            failed.set()
        with cond:
            for c, v in zip(ids, vals):
                answers[c] = float(v)
            cond.notify_all()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return np.array([r.fun for r in results]), np.stack([np.atleast_1d(r.x) for r in results])
