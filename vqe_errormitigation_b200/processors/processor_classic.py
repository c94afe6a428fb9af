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
from .. import engine
from .processor_quantum import QPU


class OptimizationTrialResult:
    """Minimal stand-in for openfermioncirq's result object (``.optimal_value`` / ``.optimal_parameters``)."""

    def __init__(self, optimal_value: float, optimal_parameters: np.ndarray, num_evaluations: int = 0):
        self.optimal_value = optimal_value
        self.optimal_parameters = optimal_parameters
        self.num_evaluations = num_evaluations


class CPU:
    @staticmethod
    def get_collection_ground_states(collection: IMeasurementCollection, cpu_iter: int, qpu_iter: int, lockstep: bool = False) -> List[IContainer]:
        """noise models x bond lengths x ``cpu_iter`` restarts (reference :18-44).  ``lockstep=True`` (not in the reference)
        keeps the objective and the optimiser - shot-based ``measure_observable``, scipy COBYLA, ``cpu_iter`` iterations from
        the ansatz's current parameters - and advances the ``cpu_iter`` restarts of a noisy wave function together
        (:meth:`get_custom_optimized_states_lockstep`): one launch per optimiser round instead of one per restart and round."""
        result = []
        for wave_function, description in collection.get_wave_functions():
            values = []
            if lockstep and isinstance(wave_function, INoiseWrapper) and hasattr(wave_function._ideal_wave_function, "sampled_energy_program"):
                funs, pars = CPU.get_custom_optimized_states_lockstep(wave_function, restarts=cpu_iter, max_cpu_iter=cpu_iter, max_qpu_iter=qpu_iter)
                for value, par in zip(funs, pars):
                    print(f"par: {par}, val: {value}")
                result.append(IContainer(m_param=wave_function.molecule_parameters, e_values=[float(v) for v in funs],
                                         m_data=wave_function.molecule, label=description))
                continue
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
    def get_custom_optimized_states_lockstep(n_w: INoiseWrapper, restarts: int, max_cpu_iter: int, max_qpu_iter: int) -> Tuple[np.ndarray, np.ndarray]:
        """``restarts`` independent runs of :meth:`get_custom_optimized_state` (reference :52-92) - the loop body of
        ``get_collection_ground_states`` :34-38 - advanced in lock-step.  Every run is scipy's own COBYLA on the sampled
        ``measure_observable`` objective from the same start point (the ansatz's current parameters; the reference deep-copies
        the wave function per restart), with its own shots: parameter row b, measurement circuit m draws from Philox stream
        b * 5 + m of a seed taken from numpy's global RNG per round, so ``np.random.seed`` still fixes the whole run.  One
        optimiser round of all runs = one launch of 5 * restarts circuits + one shot kernel
        (``SampledEnergyProgram.energies(params[B, P])``).  The shared parameter table is left untouched (the serial path
        leaves the last evaluated point in it).  Returns (fun[restarts], x[restarts, P])."""
        program = n_w.sampled_energy_program()
        names = list(n_w.operator_parameters)
        cols = [names.index(nm) for nm in program.prog.param_names]

        def batch_objective(xs: np.ndarray) -> np.ndarray:
            seed = int(np.random.randint(0, 2 ** 31 - 1))
            if not cols:               # a circuit without symbols: one sampled value serves every (identical) request
                return np.repeat(program.energies(None, max_qpu_iter, seed), len(xs))
            return program.energies(np.ascontiguousarray(xs[:, cols]), max_qpu_iter, seed)

        x0 = np.fromiter(n_w.operator_parameters.dict.values(), dtype=float)
        return lockstep_minimize(batch_objective, np.tile(x0, (int(restarts), 1)), max_cpu_iter, tol=1e-6)

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

    @staticmethod
    def get_collection_ground_states_batched(collection: IMeasurementCollection, cpu_iter: int, qpu_iter: int = 0, max_evals: int = 60,
                                             tol: float = 1e-9, x0_spread: float = 0.0, seed: int = 0) -> List[IContainer]:
        """``get_collection_ground_states`` (reference :18-44: noise models x bond lengths x ``cpu_iter`` restarts, one
        scipy run after the other) with the batch axes used: per noise model ONE plan - the noisy ansatz circuit - whose
        Hamiltonian carries one coefficient row per bond length (every H2 geometry has the same Jordan-Wigner strings),
        and (bond lengths x restarts) optimiser chains advanced in lock-step by :func:`lockstep_nelder_mead`, so every
        optimiser round is one launch over all chains.  Objective: the deterministic Tr(rho H) the reference keeps at
        :73-75 (``qpu_iter`` - shots per circuit of the sampled objective - is accepted for signature compatibility and
        unused).  Restart j starts from the ansatz's current parameters plus ``x0_spread`` * N(0, 1) (restarts of a
        deterministic objective from one point are identical).  Returns the reference's list of ``IContainer``."""
        rng = np.random.default_rng(seed)
        result: List[IContainer] = []
        for channel in (collection.noise_model_space or [None]):
            points = []         # (molecule parameters copy, molecule data, terms) per bond length
            circuit = qubits = names = x_start = label = None
            for value in collection.parameter_space:
                w = collection._bind(value)
                n_w = w if channel is None else INoiseWrapper(w, channel)
                if circuit is None:
                    label = "No Noise Model" if channel is None else channel.get_description()
                    if channel is None:
                        circuit = QPU.get_initial_state_circuit(w=w)
                        circuit.append(w.operations(w.qubits))
                    else:
                        circuit = n_w.get_noisy_circuit()
                    qubits, names = w.qubits, list(w.operator_parameters)
                    x_start = np.fromiter(w.operator_parameters.dict.values(), dtype=float)
                points.append((copy.deepcopy(w.molecule_parameters), w.molecule, QPU.get_hamiltonian_terms(w)))
            # union of the Pauli strings, one coefficient row per bond length
            strings: dict = {}
            for _, _, (xs, zs, cs) in points:
                for x, z in zip(xs, zs):
                    strings.setdefault((int(x), int(z)), len(strings))
            rows = np.zeros((len(points), len(strings)))
            for g, (_, _, (xs, zs, cs)) in enumerate(points):
                for x, z, c in zip(xs, zs, cs):
                    rows[g, strings[(int(x), int(z))]] += c
            keys = list(strings)
            terms = (np.array([k[0] for k in keys], np.uint32), np.array([k[1] for k in keys], np.uint32), rows[0].copy())
            builder, bq, measurements, _lifted, _key = engine.lower_circuit(circuit, qubits)
            prog = engine.CircuitProgram(builder, qubits=bq, measurements=measurements, hamiltonian=terms, coefficient_groups=rows)
            cols = [names.index(nm) for nm in prog.param_names]
            G, R = len(points), max(int(cpu_iter), 1)
            group_of_chain = np.repeat(np.arange(G, dtype=np.int32), R)
            x0s = np.tile(x_start, (G * R, 1))
            if x0_spread:
                x0s = x0s + x0_spread * rng.standard_normal(x0s.shape)

            def objective(xs: np.ndarray, chain_ids: np.ndarray) -> np.ndarray:
                return prog.energies_host(np.ascontiguousarray(xs[:, cols]) if cols else None, batch=len(xs),
                                          groups=np.ascontiguousarray(group_of_chain[chain_ids]))

            fun, _x, _n = lockstep_nelder_mead_grouped(objective, x0s, max_evals=max_evals, tol=tol)
            for g, (m_param, m_data, _) in enumerate(points):
                result.append(IContainer(m_param=m_param, e_values=[float(v) for v in fun[g * R:(g + 1) * R]], m_data=m_data, label=label))
        return result

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


def lockstep_nelder_mead(batch_objective: Callable[[np.ndarray], np.ndarray], x0s: np.ndarray, max_evals: int, tol: float = 1e-6,
                         initial_step: float = 0.05) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """:func:`lockstep_nelder_mead_grouped` for an objective that does not need to know which chain a row belongs to."""
    return lockstep_nelder_mead_grouped(lambda rows, _ids: batch_objective(rows), x0s, max_evals, tol, initial_step)


def lockstep_nelder_mead_grouped(batch_objective: Callable[[np.ndarray, np.ndarray], np.ndarray], x0s: np.ndarray, max_evals: int,
                                 tol: float = 1e-6, initial_step: float = 0.05) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Nelder-Mead for ``C`` independent chains, vectorised over the chains in numpy (no thread per chain): every round
    of the algorithm - reflection, then expansion / contraction where a chain asks for one, then the shrink points - is
    ONE call of ``batch_objective(rows[M, P], chain_ids[M]) -> values[M]``, i.e. one GPU launch however many chains there are (the
    engine's batch of 65 536 circuits is 65 536 chains).  ``max_evals`` bounds the objective evaluations of every chain
    (the meaning of ``max_cpu_iter`` / COBYLA's ``maxiter`` in the reference, processor_classic.py:88-92).  A chain stops
    when its simplex has collapsed to ``tol`` in both value and position.  Returns (fun[C], x[C, P], evaluations[C])."""
    x0s = np.atleast_2d(np.asarray(x0s, dtype=float))
    C, P = x0s.shape
    simplex = np.repeat(x0s[:, None, :], P + 1, axis=1)                  # [C, P+1, P]
    for i in range(P):
        step = np.maximum(initial_step * np.abs(x0s[:, i]), initial_step)      # never a degenerate (near-zero) edge
        simplex[:, i + 1, i] += step
    fvals = np.asarray(batch_objective(simplex.reshape(-1, P), np.repeat(np.arange(C), P + 1)), dtype=float).reshape(C, P + 1)
    evals = np.full(C, P + 1)
    active = np.ones(C, dtype=bool)
    idx = np.arange(C)
    while True:
        order = np.argsort(fvals, axis=1, kind="stable")
        fvals = np.take_along_axis(fvals, order, axis=1)
        simplex = np.take_along_axis(simplex, order[:, :, None], axis=1)
        spread_f = np.max(np.abs(fvals[:, 1:] - fvals[:, :1]), axis=1)
        spread_x = np.max(np.abs(simplex[:, 1:, :] - simplex[:, :1, :]), axis=(1, 2))
        active &= ~((spread_f <= tol) & (spread_x <= tol)) & (evals + 2 <= max_evals)
        if not active.any():
            break
        a = idx[active]
        centroid = simplex[a, :P, :].mean(axis=1)
        worst = simplex[a, P, :]
        xr = centroid + (centroid - worst)
        fr = np.asarray(batch_objective(xr, a), dtype=float)
        evals[a] += 1
        f_best, f_second, f_worst = fvals[a, 0], fvals[a, P - 1] if P > 1 else fvals[a, 0], fvals[a, P]
        expand = fr < f_best
        accept = ~expand & (fr < f_second)
        outside = ~expand & ~accept & (fr < f_worst)
        inside = ~expand & ~accept & ~outside
        second = expand | outside | inside
        x2 = np.where(expand[:, None], centroid + 2.0 * (xr - centroid),
                      np.where(outside[:, None], centroid + 0.5 * (xr - centroid), centroid - 0.5 * (centroid - worst)))
        f2 = np.full(len(a), np.inf)
        if second.any():
            f2[second] = np.asarray(batch_objective(x2[second], a[second]), dtype=float)
            evals[a[second]] += 1
        new_x, new_f = worst.copy(), f_worst.copy()
        take_r = accept | (expand & ~(f2 < fr)) | (outside & ~(f2 <= fr))
        new_x[take_r], new_f[take_r] = xr[take_r], fr[take_r]
        take_2 = (expand & (f2 < fr)) | (outside & (f2 <= fr)) | (inside & (f2 < f_worst))
        new_x[take_2], new_f[take_2] = x2[take_2], f2[take_2]
        simplex[a, P, :], fvals[a, P] = new_x, new_f
        shrink = (outside & ~(f2 <= fr)) | (inside & ~(f2 < f_worst))
        shrink &= evals[a] + P <= max_evals
        if shrink.any():
            sa = a[shrink]
            simplex[sa, 1:, :] = simplex[sa, :1, :] + 0.5 * (simplex[sa, 1:, :] - simplex[sa, :1, :])
            fvals[sa, 1:] = np.asarray(batch_objective(simplex[sa, 1:, :].reshape(-1, P), np.repeat(sa, P)), dtype=float).reshape(len(sa), P)
            evals[sa] += P
    best = np.argmin(fvals, axis=1)
    return fvals[idx, best], simplex[idx, best, :], evals


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
