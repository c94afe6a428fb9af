"""Host-side mirrors of the reference modules, checked on CPU.  Where an evaluation is needed the C oracle stands
in for the GPU through tests/oracle_backend.py (a test double; the product has no CPU path)."""
import numpy as np
import pytest

import oracle_backend
from oracle import dm_oracle as O
from workloads import ALPHA_STAR, h2_golden, h2_terms
from vqe_errormitigation_b200 import cirq_compat as cirq
from vqe_errormitigation_b200 import openfermion_compat as of
from vqe_errormitigation_b200.circuit_noise_extension import Noisify
from vqe_errormitigation_b200.data_containers.helper_interfaces.i_collection import IMeasurementCollection, ISimilarityCollection
from vqe_errormitigation_b200.data_containers.helper_interfaces.i_noise_wrapper import INoiseModel, INoiseWrapper
from vqe_errormitigation_b200.data_containers.helper_interfaces.i_parameter import IParameter
from vqe_errormitigation_b200.data_containers.model_hydrogen import HydrogenAnsatz
from vqe_errormitigation_b200.error_mitigation import (BasisUnitarySet, IBasisGateSingle, IErrorMitigationManager, QuasiCircuitIdentifier,
                                                       SingleQubitPauliChannel, TwoQubitDepolarizingChannel, TwoQubitPauliChannel,
                                                       reconstruct_from_basis)
from vqe_errormitigation_b200.processors.processor_classic import CPU, lockstep_minimize
from vqe_errormitigation_b200.processors.processor_quantum import QPU


# ---- cirq_compat ----------------------------------------------------------------------------------------------

def test_gate_matrices_match_oracle_definitions():
    for t in (0.3, -1.1, np.pi):
        assert np.allclose(cirq.unitary(cirq.rx(t)), O.rx(t))
        assert np.allclose(cirq.unitary(cirq.ry(t)), O.ry(t))
        assert np.allclose(cirq.unitary(cirq.rz(t)), O.rz(t))
    assert np.allclose(cirq.unitary(cirq.H), O.H) and np.array_equal(cirq.unitary(cirq.CNOT), O.CNOT)
    for gate, kraus in ((cirq.depolarize(0.1), O.depolarize(0.1)), (cirq.bit_flip(0.2), O.bit_flip(0.2)),
                        (cirq.phase_flip(0.3), O.phase_flip(0.3)), (cirq.amplitude_damp(0.1), O.amplitude_damp(0.1)),
                        (cirq.phase_damp(0.1), O.phase_damp(0.1)), (cirq.asymmetric_depolarize(0.01, 0.02, 0.03), O.asymmetric_depolarize(0.01, 0.02, 0.03)),
                        (SingleQubitPauliChannel(0.01, 0.02, 0.03), O.asymmetric_depolarize(0.01, 0.02, 0.03)),
                        (TwoQubitPauliChannel(0.01, 0.02, 0.03), O.two_qubit_pauli_channel(0.01, 0.02, 0.03)),
                        (TwoQubitDepolarizingChannel(0.1), O.two_qubit_depolarize(0.1))):
        got = cirq.channel(gate)
        assert len(got) == len(kraus) and all(np.allclose(a, b) for a, b in zip(got, kraus))


def test_circuit_earliest_strategy_value_equality_and_resolution():
    import sympy
    q = cirq.LineQubit.range(3)
    c = cirq.Circuit()
    c.append([cirq.H(q[0]), cirq.H(q[1]), cirq.CNOT(q[0], q[1]), cirq.H(q[2])])
    assert len(c.moments) == 2 and len(c.moments[0]) == 3           # H(q2) slides back to the first moment
    assert cirq.rx(0.5) == cirq.rx(0.5) and hash(cirq.depolarize(0.1)) == hash(cirq.depolarize(0.1))
    assert cirq.rx(0.5) != cirq.rx(0.6) and {cirq.H: 1}[cirq.HPowGate()] == 1
    a = sympy.Symbol("a")
    c2 = cirq.Circuit([cirq.rz(2 * a).on(q[0])])
    assert cirq.is_parameterized(c2)
    r = cirq.resolve_parameters(c2, cirq.ParamResolver({"a": 0.25}))
    assert np.allclose(cirq.unitary(next(r.all_operations())), O.rz(0.5))
    assert cirq.linear_angle(-2.0 * a + 0.5) == ("a", -2.0, 0.5)
    assert sorted([cirq.GridQubit(1, 0), cirq.GridQubit(0, 1)])[0] == cirq.GridQubit(0, 1)
    u = cirq.Circuit([cirq.H(q[0]), cirq.CNOT(q[0], q[1]), cirq.measure(q[0], key="M")]).unitary()
    assert np.allclose(u, O.CNOT @ np.kron(O.H, O.I2))


def test_cswap_decomposition_is_unitary_equivalent():
    # (control, target, target) positions -> form cirq 0.8.2 picks by adjacency, gate count
    for xs, count in (((0, 1, 2), 21), ((0, 1, 4), 24), ((0, 2, 3), 21), ((0, 3, 6), 24), ((1, 0, 2), 24), ((5, 1, 2), 21)):
        q = [cirq.LineQubit(x) for x in xs]
        ops = cirq.CSWAP(*q)._decompose_()
        assert len(ops) == count and all(len(op.qubits) <= 2 for op in ops)
        dec = cirq.Circuit(ops).unitary()
        want = cirq.Circuit([cirq.CSWAP(*q)]).unitary()
        k = np.argmax(np.abs(want))
        assert np.allclose(dec, dec.flat[k] / want.flat[k] * want)


# ---- ansatz / noise construction ----------------------------------------------------------------------------------

def test_hydrogen_ansatz_reproduces_reference_gate_list():
    """Clean circuit == the 122-gate list of SURVEY App. A.3 (oracle.h2_gate_list), op by op per qubit."""
    w = HydrogenAnsatz()
    assert list(w.operator_parameters) == ["alpha0"] and len(w.qubits) == 4
    n_w = INoiseWrapper(w, INoiseModel.empty())
    w.operator_parameters["alpha0"] = 0.123
    clean = QPU.get_resolved_circuit(n_w.get_clean_circuit(), w.operator_parameters)
    n, ops = O.from_circuit(clean)
    ref = O.h2_gate_list(0.123)
    assert n == 4 and len(ops) == len(ref) == 122
    for qb in range(4):     # EARLIEST packing may interleave disjoint qubits; per-qubit order is what matters
        mine = [(qs, m) for _, qs, m in ops if qb in qs]
        theirs = [(qs, m) for _, qs, m in ref if qb in qs]
        assert len(mine) == len(theirs)
        for (qa, ma), (qb_, mb) in zip(mine, theirs):
            assert qa == qb_ and np.allclose(ma, mb)
    ham = O.pauli_terms_to_matrix(4, *QPU.get_hamiltonian_terms(w))
    assert np.allclose(ham, O.pauli_terms_to_matrix(4, *h2_terms()))
    assert abs(O.energy_sparse(O.simulate(4, ops), ham) - O.energy_sparse(O.simulate(4, ref), ham)) < 1e-13


def test_numpy119_sign_trap():
    from vqe_errormitigation_b200.data_containers.helper_interfaces.i_wave_function import _numpy119_sign
    assert _numpy119_sign(0.0142j) == 1.0 and _numpy119_sign(-0.0142j) == -1.0 and _numpy119_sign(-2 + 5j) == -1.0


def test_noisify_and_noise_model():
    w = HydrogenAnsatz()
    model = INoiseModel([cirq.bit_flip(0.1), cirq.phase_flip(0.2)], [TwoQubitDepolarizingChannel(0.05)], "bf+pf")
    n_w = INoiseWrapper(w, model)
    ops = list(n_w.get_noisy_circuit().all_operations())
    assert len(ops) == 74 * 3 + 48 * 2             # two channels after each 1q gate, one after each CNOT
    assert n_w.operator_parameters is w.operator_parameters      # aliasing the reference relies on
    assert model.get_description() == "bf+pf" and INoiseModel.empty().get_operators(w.qubits[0]) == []
    with pytest.raises(NotImplementedError):
        model.get_operators(*w.qubits[:3])
    g = np.array([[0.3, 0.1], [0.1, 0.7]], dtype=complex)
    want = g
    for kr in (O.bit_flip(0.1), O.phase_flip(0.2)):
        want = sum(k @ want @ k.conj().T for k in kr)
    assert np.allclose(model.get_effective_gate(g), want)
    # channels without a mixture (amplitude damping) are skipped, like the swallowed TypeError in the reference
    assert np.allclose(INoiseModel([cirq.amplitude_damp(0.3)], [], "ad").get_effective_gate(g), g)
    wrapped = Noisify.wrap_noise_type([[cirq.H(w.qubits[0]), cirq.CNOT(w.qubits[0], w.qubits[1])]], model.get_operators)
    assert [type(o.gate).__name__ for o in wrapped] == ["HPowGate", "BitFlipChannel", "PhaseFlipChannel", "CNotPowGate", "TwoQubitDepolarizingChannel"]


# ---- QP error mitigation bookkeeping ---------------------------------------------------------------------------------

def test_basis_set_ptm_and_quasiprobabilities_match_oracle():
    mine = list(BasisUnitarySet.get_basis_unitary_set())
    for a, b in zip(mine, O.basis_operations()):
        assert np.allclose(a, b)
    assert len(list(BasisUnitarySet.get_basis_set(4))) == 256
    assert np.allclose(list(BasisUnitarySet.get_basis_set(4))[16 * 3 + 5], np.kron(mine[3], mine[5]))
    assert np.allclose(BasisUnitarySet.pauli_transfer_matrix(O.H), O.pauli_transfer_matrix(O.H))
    model = INoiseModel([cirq.depolarize(1e-3)], [TwoQubitDepolarizingChannel(1e-3)], "d")
    q = cirq.LineQubit.range(2)
    lookup = IErrorMitigationManager.get_qp_lookup(cirq.Circuit([cirq.H(q[0]), cirq.CNOT(q[0], q[1]), cirq.measure(q[0], key="M")]), model.get_effective_gate)
    assert np.allclose(lookup[cirq.H], O.quasi_probabilities(O.H, [O.depolarize(1e-3)]), atol=1e-14)
    assert np.allclose(lookup[cirq.CNOT], O.quasi_probabilities(O.CNOT, [O.two_qubit_depolarize(1e-3)]), atol=1e-13)
    qp = lookup[cirq.H]
    rec = reconstruct_from_basis(qp, list(BasisUnitarySet.get_ptm_basis_set(2)))
    clean = BasisUnitarySet.pauli_transfer_matrix(O.H)
    noisy = BasisUnitarySet.pauli_transfer_matrix(O.H, noise_wrapper=model.get_effective_gate)
    assert np.allclose(rec, clean @ np.linalg.inv(noisy), atol=1e-12)


def test_updated_circuit_structure_and_identifier_sampling():
    q = cirq.LineQubit.range(2)
    clean = cirq.Circuit([cirq.H(q[0]), cirq.CNOT(q[0], q[1]), cirq.measure(q[0], key="M")])
    model = INoiseModel([cirq.depolarize(0.01)], [TwoQubitDepolarizingChannel(0.01)], "d")
    circ = IErrorMitigationManager.get_updated_circuit(clean, model.get_operators, [0, 16 * 1 + 3], include_measurements=True)
    names = [type(o.gate).__name__ for o in circ.all_operations()]
    assert names == ["HPowGate", "DepolarizingChannel", "IBasisGateSingle", "CNotPowGate", "TwoQubitDepolarizingChannel",
                     "IBasisGateTwo", "TwoQubitDepolarizingChannel", "MeasurementGate"]
    two = [o for o in circ.all_operations() if type(o.gate).__name__ == "IBasisGateTwo"][0]
    b = O.basis_operations()
    assert np.allclose(cirq.unitary(two), np.kron(b[1], b[3]))
    mgr = IErrorMitigationManager(clean, model, None)
    np.random.seed(3)
    mgr.set_identifier_range(5000)
    assert mgr._counts.sum() == 5000 and mgr._rows.shape[1] == 2
    assert set(np.unique(mgr._rows[:, 0])) <= {0, 1, 2, 3}                      # only Pauli corrections are ever drawn
    frac_identity = mgr._counts[(mgr._rows == 0).all(axis=1)].sum() / 5000
    assert abs(frac_identity - (1 - 0.0132) * (1 - 0.0105)) < 0.02
    ident = QuasiCircuitIdentifier(clean, mgr._gate_lookup)
    assert len(ident.identifiers_id) == 2 and len(ident.identifiers_weight) == 3 and ident.sign in (-1, 1)
    d = mgr._dict_identifiers
    assert sum(d.values()) == 5000 and all(isinstance(k, QuasiCircuitIdentifier) for k in d)


def test_mitigated_expectation_recovers_noiseless_value(monkeypatch):
    """Density-representation QP mitigation on H2: mean over sampled variants -> noiseless energy within the
    sampling error, and exactly the weighted oracle mean of the same variant table."""
    oracle_backend.install(monkeypatch)
    w = HydrogenAnsatz()
    w.operator_parameters["alpha0"] = ALPHA_STAR
    model = INoiseModel([cirq.depolarize(1e-3)], [TwoQubitDepolarizingChannel(1e-3)], "d")
    n_w = INoiseWrapper(w, model)
    clean = QPU.get_resolved_circuit(n_w.get_clean_circuit(), w.operator_parameters)
    ham = QPU.get_hamiltonian_objective_operator(w)
    mgr = IErrorMitigationManager(clean, model, ham)
    np.random.seed(11)
    mgr.set_identifier_range(4000)
    values, mu = mgr.get_mu_effective(error_mitigation=True, density_representation=True)
    assert len(values) == 4000
    fci = h2_golden()["molecules"]["0.7414"]["fci_energy"]
    # exact cross-check: same rows through the oracle's own variant construction (manager state is still
    # "mitigation on": like the reference, get_mu_effective leaves its switches set)
    gates = O.h2_gate_list(ALPHA_STAR)
    hm = O.pauli_terms_to_matrix(4, *h2_terms())
    signs, weights = mgr._signs_and_weights(mgr._rows)
    total = 0.0
    for row, cnt, sg in list(zip(mgr._rows, mgr._counts, signs))[:12]:
        # identifiers index gates in clean_circuit.all_operations() order (moment-packed), so the independent
        # evaluation goes through get_updated_circuit (the reference-shaped per-variant circuit) + the numpy oracle
        variant_circuit = IErrorMitigationManager.get_updated_circuit(clean, model.get_operators, [int(v) for v in row])
        e = O.energy_sparse(O.simulate(4, O.from_circuit(variant_circuit, w.qubits)[1]), hm)
        got = mgr.process_circuit((QuasiCircuitIdentifier.from_arrays(row, sg, weights), int(cnt)))
        assert len(got) == cnt and abs(got[0] - sg * np.prod(weights) * e) < 1e-10
    _, noisy = mgr.get_mu_effective(error_mitigation=False, density_representation=True)
    assert abs(noisy - (-1.0559477031919418)) < 1e-10          # golden noisy energy at alpha*
    assert abs(mu - fci) < 0.25 * abs(noisy - fci)               # mitigation removes most of the bias


def test_observable_conventions(monkeypatch):
    """ndarray objective -> element-wise product (only diag(H) counts); None -> Z on the LAST qubit."""
    oracle_backend.install(monkeypatch)
    q = cirq.LineQubit.range(2)
    clean = cirq.Circuit([cirq.rx(0.7).on(q[0]), cirq.ry(1.1).on(q[1]), cirq.CNOT(q[0], q[1])])
    model = INoiseModel.empty()
    rho = O.simulate(2, O.from_circuit(clean)[1])
    for objective, want in ((None, O.energy_elementwise(rho, np.kron(np.eye(2), O.Z))),
                            (np.kron(O.X, O.Z) + np.kron(O.Z, O.I2), O.energy_elementwise(rho, np.kron(O.X, O.Z) + np.kron(O.Z, O.I2)))):
        mgr = IErrorMitigationManager(clean, model, objective)
        mgr.set_identifier_range(3)
        _, mu = mgr.get_mu_effective(error_mitigation=True, density_representation=True)
        assert abs(mu - want) < 1e-12


# ---- optimiser drivers --------------------------------------------------------------------------------------------------

def test_qpu_energy_and_cobyla_paths(monkeypatch):
    oracle_backend.install(monkeypatch)
    w = HydrogenAnsatz()
    model = INoiseModel([cirq.depolarize(1e-3)], [TwoQubitDepolarizingChannel(1e-3)], "d")
    n_w = INoiseWrapper(w, model)
    circuit = n_w.get_noisy_circuit()
    e = QPU.get_simulated_noisy_expectation_value(n_w, circuit, cirq.ParamResolver({"alpha0": ALPHA_STAR}))
    assert abs(e - (-1.0559477031919418)) < 1e-10
    es = CPU.get_batched_energies(n_w, np.array([[0.0], [ALPHA_STAR]]))
    assert np.allclose(es, [-1.0380841518615922, -1.0559477031919418], atol=1e-10)
    res = CPU.get_optimized_state(w, max_iter=60)
    assert abs(res.optimal_value - h2_golden()["molecules"]["0.7414"]["fci_energy"]) < 1e-6
    fun, x = CPU.get_custom_optimized_state(n_w, max_cpu_iter=40, max_qpu_iter=10, use_density=True)
    assert fun < -1.0559 and abs(x[0] - ALPHA_STAR) < 5e-3
    assert abs(w.operator_parameters["alpha0"] - x[0]) < 1e-2     # last evaluated value stays in the shared dict


def test_lockstep_minimize_matches_single_chain():
    from scipy import optimize
    calls = []

    def batch(xs):
        calls.append(len(xs))
        return np.sum((xs - np.array([0.3, -0.2])) ** 2, axis=1) + 1.0

    x0s = np.array([[0.0, 0.0], [1.0, 1.0], [-0.5, 0.4]])
    funs, xs = lockstep_minimize(batch, x0s, maxiter=60)
    for c in range(3):
        ref = optimize.minimize(lambda x: float(np.sum((x - np.array([0.3, -0.2])) ** 2) + 1.0), x0s[c], method="COBYLA",
                                options={"maxiter": 60}, tol=1e-6)
        assert abs(funs[c] - ref.fun) < 1e-12 and np.allclose(xs[c], ref.x)
    assert max(calls) == 3                                       # rounds are evaluated as one batch


def test_containers_and_parameters():
    p = IParameter({"a": 1.0, "b": 2.0})
    p.update([3.0, 4.0])
    assert p["a"] == 3.0 and list(p) == ["a", "b"] and (p + IParameter({"c": 1}))["c"] == 1
    with pytest.raises(IndexError):
        p.update([1.0])
    assert ISimilarityCollection.get_settings(1e-4, 100, 10000) == "P0.0001_R100_C10000"
    w = HydrogenAnsatz()
    coll = IMeasurementCollection(w, [0.7, 0.8], [INoiseModel.empty()])
    seen = [(wf.molecule_parameters["r0"], desc) for wf, desc in coll.get_wave_functions()]
    assert seen == [(0.7, "(p=0)"), (0.8, "(p=0)")]
    assert abs(w.molecule.fci_energy - h2_golden()["molecules"]["0.8"]["fci_energy"]) < 1e-15


def test_molecular_data_from_reference_hdf5_if_present():
    import os
    path = "/root/reference/src/mol_data/H2_0.7414"
    if not os.path.exists(path + ".hdf5"):
        pytest.skip("reference checkout not present (GPU box)")
    m = of.MolecularData(description="0.7414", filename=path).load()
    rec = h2_golden()["molecules"]["0.7414"]
    assert m.hf_energy == rec["hf_energy"] and m.fci_energy == rec["fci_energy"] and m.n_qubits == 4
    packaged = of.MolecularData(description="0.7414", filename="/nonexistent/H2_0.7414").load()
    assert np.array_equal(m.two_body_integrals, packaged.two_body_integrals)
    assert np.array_equal(m.ccsd_double_amps, packaged.ccsd_double_amps)


def test_experiment_and_demo_drivers_on_the_oracle_backend(monkeypatch, tmp_path, capsys):
    """The compute halves of visual_testing.py's similarity experiments, main.custom_ansatz / get_log_experiment and the
    calculate_minimisation driver, run end to end on CPU with the oracle standing in for the engine (reduced budgets)."""
    from vqe_errormitigation_b200 import calculate_minimisation as CM
    from vqe_errormitigation_b200 import experiments as E
    from vqe_errormitigation_b200 import main as M
    oracle_backend.install(monkeypatch)
    np.random.seed(3)
    fci = h2_golden()["molecules"]["0.7414"]["fci_energy"]
    d = E.similarity_hydrogen_circuit(prob=1e-4, measure_count=2, identifier_count=2000, fixed_parameters=[ALPHA_STAR])
    assert d._settings == "P0.0001_R2_C2000" and d.mu_ideal == -1.13727 and len(d.data_noisy.get_data) == 2
    # 2000 // 5 identifiers: noisy ~ -1.031 (SURVEY App. A.4 row 3 plus measurement-basis noise), mitigated ~ FCI
    assert abs(d.data_noisy.get_mean - (-1.031)) < 0.03 and abs(d.data_mitigated.get_mean - fci) < 0.08
    monkeypatch.setattr(CM, "DATA_DIR", str(tmp_path / "classic_minimisation"))
    CM.write_object(d, "H2_Similarity_" + d._settings)
    back = CM.read_object("H2_Similarity_" + d._settings)
    assert back.data_mitigated.get_data == d.data_mitigated.get_data and back.mu_ideal == d.mu_ideal
    # main.custom_ansatz: bit + phase flip model, shot-based COBYLA (reference main.py:35-50) on a reduced budget
    value, params = M.custom_ansatz(max_cpu_iter=3, max_qpu_iter=200, p=1e-3)
    assert -1.3 < value < -0.9 and len(params) == 1
    errors = M.get_log_experiment([10, 1000], expectation=0.5, experiment=lambda shots: 0.5 + 1.0 / shots, reps=2)
    assert errors == {10: pytest.approx(0.1), 1000: pytest.approx(1e-3)}
    capsys.readouterr()


def test_remaining_experiment_loops_on_the_oracle_backend(monkeypatch, capsys):
    """visual_testing.py's other experiment loops (density-vs-shots, repeated noisy optimisation, raw-vs-mitigated per noise
    level) at toy budgets.  full_raw_vs_mitigation_per_noise hands an INoiseWrapper to the manager as its objective
    (reference :761), which must be accepted like an IWaveFunction."""
    from vqe_errormitigation_b200 import experiments as E
    oracle_backend.install(monkeypatch)
    np.random.seed(2)
    err, spread, probs, reps = E.hamiltonian_density_vs_measure(prob_list=[1e-4], meas_reps=[10, 4000], count=6, max_iter=40)
    assert err.shape == spread.shape == (1, 2) and probs == [1e-4] and reps == [10, 4000]
    assert spread[0, 1] < spread[0, 0] and err[0, 1] < 0.02          # more shots: tighter, and close to Tr(rho H)
    exp_list, opt_list, probs = E.circuit_parameter_optimization(prob_list=[1e-4], count=2, max_cpu_iter=10, max_qpu_iter=4000)
    assert len(exp_list) == len(opt_list) == 1 and len(exp_list[0]) == 2 and all(-1.1 < v < -0.95 for v in exp_list[0])
    assert all(abs(a) < 0.2 for a in opt_list[0])                    # COBYLA's halving steps end near alpha* = 0.0141
    noisy, mitigated, probs = E.full_raw_vs_mitigation_per_noise(measure_count=1, identifier_count=2000, prob_list=[1e-4], max_cpu_iter=6)
    pairs = E.raw_and_mitigated_repeated(INoiseWrapper(HydrogenAnsatz(), E.asymmetric_pauli_noise_model(1e-4)), E.asymmetric_pauli_noise_model(1e-4),
                                         repetitions=2, max_qpu_iter=2000, max_cpu_iter=6)
    assert len(pairs) == 2 and all(-1.2 < a < -0.9 and -1.3 < b < -0.9 for a, b in pairs)
    fci = h2_golden()["molecules"]["0.7414"]["fci_energy"]
    assert len(noisy) == len(mitigated) == 1 and abs(mitigated[0][0] - fci) < 0.12 and -1.1 < noisy[0][0] < -0.9
    # bond-length sweeps: r = 0.0 has no molecular data and H2_1.8.hdf5 stops after the SCF step -> NaN rows (reference :983-986)
    r, fci_e, hf_e, measured = E.hydrogen_fci_energy(atomic_distance_list=[0.0, 0.7, 1.8], repetitions=2, max_cpu_iter=8, max_qpu_iter=2000)
    assert np.isnan(fci_e[0]) and np.isnan(fci_e[2]) and np.all(np.isnan(measured[0])) and np.all(np.isnan(measured[2]))
    assert abs(fci_e[1] - h2_golden()["molecules"]["0.7"]["fci_energy"]) < 1e-12 and all(-1.2 < v < -0.9 for v in measured[1])
    first = E.hydrogen_energy_spectrum(measure_count=1, identifier_count=1000, atomic_distance_list=[0.4, 1.0], max_cpu_iter=6)
    more = E.hydrogen_energy_spectrum(measure_count=1, identifier_count=1000, atomic_distance_list=[0.4, 1.0], previous=first, max_cpu_iter=6)
    assert [len(v) for v in more[0]] == [2, 2] and more[0][0][0] == first[0][0][0]          # resumed run extends the stored lists
    assert abs(more[3][1] - h2_golden()["molecules"]["1.0"]["fci_energy"]) < 1e-12
    errors, fit = E.sampling_noise_scaling(shot_list=[20, 2000], reps=3, prob=1e-4, max_iter=60)
    assert list(errors) == [20, 2000] and errors[2000] < errors[20] < 1.0 and fit is not None and fit[0] < 0
    assert errors[2000] < 0.03           # the sparse objective is Tr(rho H) with the full H: the estimate converges on the optimum
    q = cirq.LineQubit.range(2)
    rho = E.simulate_density_matrix(cirq.Circuit([cirq.H(q[0]), cirq.measure(q[0], key="a"), cirq.CNOT(q[0], q[1])]))
    assert np.allclose(rho, np.diag([0.5, 0, 0, 0.5]), atol=1e-12)          # measurement = dephasing: no coherence survives
    capsys.readouterr()


def test_lockstep_nelder_mead_vectorised_over_chains():
    """Every round of the batched Nelder-Mead is one objective call over all chains (no thread per chain): thousands of
    chains cost a few hundred calls, each chain converges to the minimiser scipy finds, and the grouped form hands the
    objective the chain ids."""
    from scipy.optimize import minimize
    from vqe_errormitigation_b200.processors.processor_classic import lockstep_nelder_mead, lockstep_nelder_mead_grouped
    calls = [0]

    def f(xs):
        calls[0] += 1
        return ((xs - np.array([1.5, -0.3])) ** 2).sum(axis=1) + 0.05 * np.sin(xs[:, 0]) ** 2

    rng = np.random.default_rng(0)
    x0 = rng.normal(scale=0.3, size=(3000, 2))
    fun, x, n = lockstep_nelder_mead(f, x0, max_evals=400, tol=1e-10)
    ref = minimize(lambda v: f(v[None, :])[0], np.zeros(2), method="Nelder-Mead", options={"xatol": 1e-10, "fatol": 1e-10})
    assert calls[0] < 1000 and np.max(np.abs(fun - ref.fun)) < 1e-9 and np.max(np.abs(x - ref.x)) < 1e-4
    assert n.max() <= 400 and n.min() > 20
    # the evaluation budget is respected (the reference's max_cpu_iter = COBYLA maxiter, processor_classic.py:88-92)
    fun10, _x, n10 = lockstep_nelder_mead(f, x0[:50], max_evals=10, tol=1e-12)
    assert n10.max() <= 10 and np.all(fun10 <= f(x0[:50]) + 1e-15)
    # one-dimensional chains with a per-chain target (the objective receives the chain ids)
    target = np.linspace(-1, 1, 9)
    fun1, x1, _ = lockstep_nelder_mead_grouped(lambda xs, ids: (xs[:, 0] - target[ids]) ** 2 + ids, np.zeros((9, 1)), 120, tol=1e-13)
    assert np.max(np.abs(x1[:, 0] - target)) < 1e-6 and np.max(np.abs(fun1 - np.arange(9))) < 1e-12


def test_lockstep_minimize_propagates_objective_failure():
    """A batched objective that raises must not leave the chain threads blocked."""
    import threading
    from vqe_errormitigation_b200.processors.processor_classic import lockstep_minimize
    calls = [0]

    def bad(xs):
        calls[0] += 1
        if calls[0] == 3:
            raise RuntimeError("boom")
        return (xs ** 2).sum(axis=1)

    before = threading.active_count()
    with pytest.raises(RuntimeError, match="boom"):
        lockstep_minimize(bad, np.ones((4, 2)), 50)
    assert threading.active_count() == before


def test_coefficient_groups_lowering():
    """dmx_plan_set_hamiltonian_groups: host-only checks - shape validation, call order, and the plan still finalises on the
    frame path with the union of the rows' non-zero pattern."""
    from workloads import h2_ops, h2_terms
    from oracle import dm_oracle as O
    from vqe_errormitigation_b200 import engine
    from vqe_errormitigation_b200._lib import DmxError
    b = engine.ProgramBuilder(4)
    for kind, qs, pl in h2_ops([O.depolarize(1e-3)], [O.two_qubit_depolarize(1e-3)]):
        if kind == "u":
            b.add_unitary(qs, pl)
        elif kind == "k":
            b.add_kraus(qs, pl)
        else:
            b.add_rotation(qs[0], pl[0], pl[1], pl[2], pl[3])
    terms = h2_terms()
    rows = np.stack([terms[2], 2.0 * terms[2], np.zeros_like(terms[2])])
    rows[2, 0] = 1.0
    prog = engine.CircuitProgram(b, hamiltonian=terms, coefficient_groups=rows)
    assert prog.n_groups == 3 and prog.info["path"] == 1 and prog.info["n_terms"] == len(terms[2])
    with pytest.raises(ValueError):
        engine.CircuitProgram(b, hamiltonian=terms, coefficient_groups=rows[:, :5])
    lib = engine._lib.load()
    import ctypes as C
    h = C.c_void_p()
    ops = (engine.DmxOp * 1)()
    assert lib.dmx_plan_create(1, 0, 0, ops, 0, None, 0, C.byref(h)) == 0
    c = np.ones(1)
    assert lib.dmx_plan_set_hamiltonian_groups(h, 1, c.ctypes.data_as(C.POINTER(C.c_double))) == -5      # strings first (DMX_ERR_STATE)
    lib.dmx_plan_destroy(h)


def test_error_mitigation_module_helpers(monkeypatch, capsys):
    """Module-level names of src/error_mitigation.py that sit beside the manager (:585-744) and the diagram symbols."""
    from vqe_errormitigation_b200.error_mitigation import (SingleQubitLeakageChannel, decomposition_time_test, get_hardcoded_noise,
                                                           simulate_error_mitigation)
    oracle_backend.install(monkeypatch)
    # matrix-level noise wrappers: trace preserving, and equal to the channel of the same name applied to the gate matrix
    g = cirq.unitary(cirq.rx(0.3))
    dep = get_hardcoded_noise("depolar", 0.03)(g)
    want = sum(p * (k @ g @ k.conj().T) for p, k in cirq.mixture(cirq.depolarize(0.03)))
    assert np.allclose(dep, want, atol=1e-15)
    g2 = cirq.unitary(cirq.CNOT)
    dep2 = get_hardcoded_noise("depolar", 0.03)(g2)
    want2 = sum(p * (k @ g2 @ k.conj().T) for p, k in cirq.mixture(TwoQubitDepolarizingChannel(0.03)))
    assert np.allclose(dep2, want2, atol=1e-15)
    z = np.diag([1.0, -1.0])
    assert np.allclose(get_hardcoded_noise("dephase", 0.2)(g), 0.8 * g + 0.2 * z @ g @ z)
    assert get_hardcoded_noise("default", 0.2)(g) is g
    with pytest.raises(NotImplementedError):
        get_hardcoded_noise("other", 0.1)
    decomposition_time_test(cirq.unitary(cirq.H))
    out = capsys.readouterr().out
    assert "Sum of quasi-probabilities: 1.0" in out.replace("0.9999999999999", "1.0") and "difference: " in out
    assert float(out.rsplit("difference: ", 1)[1]) < 1e-12
    # four-way comparison on a 2-qubit circuit with an ndarray objective: mitigation moves the noisy value back to the ideal
    q = cirq.LineQubit.range(2)
    clean = cirq.Circuit([cirq.ry(0.9).on(q[0]), cirq.CNOT(q[0], q[1]), cirq.rx(0.4).on(q[1])])
    model = INoiseModel([cirq.depolarize(0.02)], [TwoQubitDepolarizingChannel(0.02)], "d")
    objective = np.kron(O.Z, O.Z) + 0.5 * np.kron(O.Z, O.I2)
    np.random.seed(5)
    mu_ideal, mu_noisy, mu_eff = simulate_error_mitigation(clean, model, 4000, "two qubits", objective)
    rho = O.simulate(2, O.from_circuit(clean)[1])
    assert abs(mu_ideal - O.energy_elementwise(rho, objective)) < 1e-12
    assert abs(mu_noisy - mu_ideal) > 0.02 and abs(mu_eff - mu_ideal) < 0.35 * abs(mu_noisy - mu_ideal)
    # diagram symbols
    info = TwoQubitDepolarizingChannel(0.03)._circuit_diagram_info_(None)
    assert info.wire_symbols == ("D(0.002)", "D(0.002)")
    assert SingleQubitPauliChannel(0.1, 0.0, 0.0)._circuit_diagram_info_(None).wire_symbols == ("D(0.900)",)
    assert SingleQubitLeakageChannel(0.25)._circuit_diagram_info_(None).wire_symbols == ("L(0.250)",)


def test_lockstep_restarts_keep_the_sampled_cobyla_semantics(monkeypatch, capsys):
    """``get_collection_ground_states(..., lockstep=True)``: the cpu_iter restarts of a noisy wave function as lock-step COBYLA
    chains on the batched ``measure_observable`` objective - same start point, own shots per chain, values distributed like the
    serial loop's (reference processor_classic.py:18-44)."""
    oracle_backend.install(monkeypatch)
    model = INoiseModel([cirq.depolarize(1e-3)], [TwoQubitDepolarizingChannel(1e-3)], "d")
    n_w = INoiseWrapper(HydrogenAnsatz(), model)
    program = n_w.sampled_energy_program()
    assert program.prog.param_names == ["alpha0"] and program.variant_rows.shape == (5, 4)
    np.random.seed(7)
    funs, xs = CPU.get_custom_optimized_states_lockstep(n_w, restarts=6, max_cpu_iter=12, max_qpu_iter=20000)
    assert funs.shape == (6,) and xs.shape == (6, 1) and len(set(funs.tolist())) == 6          # independent shots per chain
    assert n_w.operator_parameters["alpha0"] == 0.0                                            # the shared table is untouched
    np.random.seed(7)
    again, _ = CPU.get_custom_optimized_states_lockstep(n_w, restarts=6, max_cpu_iter=12, max_qpu_iter=20000)
    assert np.array_equal(funs, again)                                                         # np.random.seed fixes the run
    serial = np.array([CPU.get_custom_optimized_state(INoiseWrapper(HydrogenAnsatz(), model), max_cpu_iter=12, max_qpu_iter=20000)[0]
                       for _ in range(6)])
    exact = -1.0559477031919418             # Tr(rho H) at alpha*: the sampled optimum scatters around it (shot noise ~ 0.6 / sqrt(shots))
    assert abs(funs.mean() - exact) < 0.02 and abs(serial.mean() - exact) < 0.02
    assert 0.25 < (funs.std() + 1e-4) / (serial.std() + 1e-4) < 4.0
    assert np.all(np.abs(xs[:, 0] - ALPHA_STAR) < 0.3)
    # through the collection driver: one container per (noise model, bond length) with cpu_iter values each
    collection = IMeasurementCollection(w=HydrogenAnsatz(), p_space=[0.7, 0.8], n_space=[model])
    containers = CPU.get_collection_ground_states(collection, cpu_iter=3, qpu_iter=4000, lockstep=True)
    assert len(containers) == 2 and all(len(c._optimized_exp_values) == 3 for c in containers)
    for c, r in zip(containers, ("0.7", "0.8")):
        assert all(abs(v - h2_golden()["molecules"][r]["fci_energy"]) < 0.15 for v in c._optimized_exp_values)
        assert abs(c.fci_value - h2_golden()["molecules"][r]["fci_energy"]) < 1e-12
    assert capsys.readouterr().out.count("par: ") == 6


def test_three_qubit_gates_lower_through_their_gate_decomposition(monkeypatch):
    """Found by tools/fuzz_lowering.py --circuits: an undecomposed TOFFOLI, or a CSWAP operation rebuilt by
    ``resolve_parameters`` (a plain GateOperation), must lower through the gate's own ``_decompose_`` - exact for the unitary."""
    import sympy
    oracle_backend.install(monkeypatch)
    q = cirq.LineQubit.range(4)
    circuit = cirq.Circuit([cirq.H(q[0]), cirq.ry(0.7)(q[1]), cirq.rx(sympy.Symbol("t"))(q[2]), cirq.TOFFOLI(q[0], q[1], q[3]),
                            cirq.CSWAP(q[3], q[0], q[2]), cirq.depolarize(0.02)(q[2])])
    resolver = cirq.ParamResolver({"t": 0.4})
    want = O.simulate(4, O.from_circuit(cirq.resolve_parameters(circuit, resolver), q)[1])
    for program, res in ((circuit, resolver), (cirq.resolve_parameters(circuit, resolver), None)):
        got = cirq.DensityMatrixSimulator().simulate(program, param_resolver=res, qubit_order=q).final_density_matrix
        assert np.abs(got - want).max() < 1e-12


def test_fermion_algebra_against_explicit_matrices():
    """openfermion_compat (the reference's ansatz and Hamiltonian builders call these, i_wave_function.py:98-139,
    processor_quantum.py:74-78): jordan_wigner, normal_ordered and hermitian_conjugated of random fermion operators against
    explicit 2^n x 2^n ladder matrices a_j = Z..Z (x) |0><1| (x) I..I, and pauli_masks against get_sparse_operator."""
    rng = np.random.default_rng(5)
    n = 4
    lower = np.array([[0, 1], [0, 0]], dtype=complex)

    def ladder(mode, dagger):
        m = np.eye(1, dtype=complex)
        for j in range(n):
            m = np.kron(m, O.Z if j < mode else lower if j == mode else O.I2)
        return m.conj().T if dagger else m

    def matrix(op):
        total = np.zeros((1 << n, 1 << n), dtype=complex)
        for term, coeff in op.terms.items():
            m = np.eye(1 << n, dtype=complex)
            for mode, dagger in term:
                m = m @ ladder(mode, dagger)
            total += coeff * m
        return total

    for _ in range(40):
        op = of.FermionOperator()
        for _ in range(int(rng.integers(1, 5))):
            term = " ".join(f"{int(rng.integers(n))}{'^' if rng.random() < 0.5 else ''}" for _ in range(int(rng.integers(1, 5))))
            op += of.FermionOperator(term, complex(rng.normal(), rng.normal()))
        want = matrix(op)
        jw = of.jordan_wigner(op)
        assert np.abs(of.get_sparse_operator(jw, n_qubits=n).toarray() - want).max() < 1e-12
        assert np.abs(matrix(of.normal_ordered(op)) - want).max() < 1e-12
        assert np.abs(matrix(of.hermitian_conjugated(op)) - want.conj().T).max() < 1e-12
        herm = of.jordan_wigner(op + of.hermitian_conjugated(op))
        x, z, c = of.pauli_masks(herm, n)
        assert np.abs(O.pauli_terms_to_matrix(n, x, z, c) - (want + want.conj().T)).max() < 1e-12


def test_unmitigated_estimate_keeps_the_reference_multiplicity(monkeypatch):
    """error_mitigation=False: the reference hands process_circuit ONE identity item of multiplicity count * meas_reps
    (error_mitigation.py:434), so meas_reps enters the shot count twice (:386) and the result list has count * meas_reps
    entries (:399).  Kept as is - all of the reference's callers use meas_reps = 1."""
    oracle_backend.install(monkeypatch)
    q = cirq.LineQubit.range(2)
    clean = cirq.Circuit([cirq.ry(0.8).on(q[0]), cirq.CNOT(q[0], q[1]), cirq.measure(q[0], key="M")])
    mgr = IErrorMitigationManager(clean, INoiseModel([cirq.depolarize(0.01)], [TwoQubitDepolarizingChannel(0.01)], "d"), None)
    mgr.set_identifier_range(50)
    seen = []
    monkeypatch.setattr("vqe_errormitigation_b200.engine.shot_expectations",
                        lambda probs, shots, masks, seed, first_row=0: (seen.append(np.asarray(shots).copy()),
                                                                        oracle_backend._shot_expectations(probs, shots, masks, seed, first_row))[1])
    values, mu = mgr.get_mu_effective(error_mitigation=False, density_representation=False, meas_reps=3)
    assert len(values) == 150 and np.all(values == values[0]) and abs(mu - values[0]) < 1e-15
    assert [int(np.sum(s)) for s in seen] == [450]


def test_circuit_text_diagram():
    """``print(circuit)`` (reference main.py:50) draws a cirq-style diagram; the reference's channel classes supply their
    own wire symbols through ``_circuit_diagram_info_`` (error_mitigation.py:784-786,821-822)."""
    import sympy
    q = cirq.LineQubit.range(3)
    circuit = cirq.Circuit([cirq.H(q[0]), cirq.rx(sympy.Symbol("a") * 2)(q[1]), cirq.CNOT(q[0], q[2]), TwoQubitDepolarizingChannel(0.015)(q[0], q[2]),
                            SingleQubitPauliChannel(0.1, 0.0, 0.0)(q[1]), cirq.CSWAP(q[0], q[1], q[2]), cirq.measure(q[0], q[2], key="M")])
    text = str(circuit)
    lines = text.splitlines()
    assert len(lines) == 5 and lines[0].startswith("0: ───H") and lines[2].startswith("1: ───Rx(") and lines[4].startswith("2: ───")
    assert "D(0.001)" in lines[0] and "D(0.001)" in lines[4] and "D(0.900)" in lines[2]          # the channels' own symbols
    assert "@" in lines[0] and "X" in lines[4] and "M('M')" in lines[0] and lines[4].rstrip("─").endswith("M")
    assert "│" in lines[1] and "│" in lines[3] and "┼" in lines[2]                                # CNOT / measure cross qubit 1
    assert len({len(line) for line in (lines[0], lines[2], lines[4])}) == 1                        # wires end together
    assert repr(circuit).splitlines()[0].strip().startswith("0: H(0)")                            # the moment list stays available
    assert str(cirq.Circuit()) == ""
