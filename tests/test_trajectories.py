"""Mid-circuit measurement (trajectories.TrajectoryProgram behind DensityMatrixSimulator.run / simulate) against the oracle's
restatement of cirq's per-repetition simulation (oracle/dm_oracle.py: run_repeat / measurement_distribution).

Every test runs twice: on the CPU test double (host logic; the variant batches go through the C oracle) and, marked ``gpu``,
on the real engine through the C ABI.  Exact parity: the outcome-tree probabilities and the collapsed final states at 1e-12
against the oracle's explicit projections of the complex density matrix.  Statistical parity: sampled frequencies within 5 sigma."""
import numpy as np
import pytest

import oracle_backend
from oracle import dm_oracle as O
from vqe_errormitigation_b200 import cirq_compat as cirq
from vqe_errormitigation_b200 import engine, trajectories
from vqe_errormitigation_b200.error_mitigation import IBasisGateSingle


@pytest.fixture(params=["oracle_double", pytest.param("gpu", marks=pytest.mark.gpu)])
def backend(request, monkeypatch):
    if request.param == "oracle_double":
        oracle_backend.install(monkeypatch)
    else:
        import torch
        assert torch.cuda.is_available()
    trajectories._CACHE.clear()
    engine._PROGRAM_CACHE.clear()
    yield request.param
    trajectories._CACHE.clear()
    engine._PROGRAM_CACHE.clear()


def _noise(circuit_ops, p=0.02):
    out = []
    for op in circuit_ops:
        out.append(op)
        if not isinstance(op.gate, cirq.MeasurementGate):
            out.extend(cirq.depolarize(p).on(q) for q in op.qubits)
    return out


def circuit_reused_qubit():
    """3 qubits: entangle, measure q0 mid-way, keep using it, measure the rest at the end (one key per gate)."""
    q = cirq.LineQubit.range(3)
    ops = [cirq.H(q[0]), cirq.CNOT(q[0], q[1]), cirq.ry(0.7)(q[2]), cirq.measure(q[0], key="a"),
           cirq.CNOT(q[1], q[2]), cirq.rx(0.4)(q[0]), cirq.CNOT(q[0], q[2]), cirq.measure(q[1], q[2], key="b"),
           cirq.measure(q[0], key="c")]
    return cirq.Circuit(_noise(ops)), q


def circuit_non_trace_preserving():
    """A projector-like 'gate' (IBasisGateSingle wraps any 2x2 matrix, error_mitigation.py:277-300) before AND after a
    measurement: the conditional probabilities must be taken on the state at the measurement, renormalised like cirq's."""
    q = cirq.LineQubit.range(2)
    damp = IBasisGateSingle(np.array([[1.0, 0.0], [0.0, 0.5]], complex))
    ops = [cirq.H(q[0]), cirq.CNOT(q[0], q[1]), damp.on(q[1]), cirq.measure(q[0], key="m0"), cirq.H(q[0]), damp.on(q[0]),
           cirq.ry(1.1)(q[1]), cirq.measure(q[0], q[1], key="end")]
    return cirq.Circuit(ops), q


def circuit_wide_measurement():
    """5 qubits, a three-qubit measurement with an invert mask in the middle, a repeated measurement of one qubit."""
    q = cirq.LineQubit.range(5)
    ops = [cirq.H(q[0]), cirq.ry(0.9)(q[1]), cirq.CNOT(q[0], q[2]), cirq.CNOT(q[1], q[3]), cirq.rx(0.3)(q[4]),
           cirq.CNOT(q[3], q[4]), cirq.measure(q[2], q[3], q[0], key="mid", invert_mask=(True, False, False)),
           cirq.measure(q[2], key="again"), cirq.CNOT(q[2], q[4]), cirq.ry(0.5)(q[0]), cirq.CNOT(q[0], q[1]),
           cirq.measure(q[4], key="late"), cirq.H(q[4]), cirq.measure(*q, key="all")]
    return cirq.Circuit(_noise(ops, 0.01)), q


CIRCUITS = {"reused": circuit_reused_qubit, "non_tp": circuit_non_trace_preserving, "wide": circuit_wide_measurement}


def _engine_distribution(tp):
    """{all measured bits in circuit order: probability} from the outcome tree + the final diagonals."""
    codes, prob = tp.joint_distribution()
    n = tp.n_qubits
    final_slots = [s[0] for s in tp.final_sites]
    out = {}
    fin = tp.final_probabilities(codes) if tp.final_sites else None
    for r in range(len(prob)):
        base = [int(c) - 1 for c in codes[r]]
        if fin is None:
            out[tuple(base)] = out.get(tuple(base), 0.0) + prob[r]
            continue
        for x in np.flatnonzero(fin[r] > 0):
            bits = list(base)
            for slot, qpos, *_ in tp.final_sites:
                bits[slot] = (int(x) >> (n - 1 - qpos)) & 1
            key = tuple(bits)
            out[key] = out.get(key, 0.0) + prob[r] * fin[r, x]
    assert set(final_slots).isdisjoint(s[0] for s in tp.collapse_sites)
    return out


def _raw_bits(measurements_spec, bits):
    """Oracle records are raw outcomes; the engine's records carry the invert mask.  bits: tuple in circuit order."""
    out, i = [], 0
    for _, positions, invert in measurements_spec:
        for j in range(len(positions)):
            out.append(bits[i] ^ (1 if j < len(invert) and invert[j] else 0))
            i += 1
    return tuple(out)


@pytest.mark.parametrize("name", sorted(CIRCUITS))
def test_outcome_tree_matches_the_oracle(backend, name):
    circuit, q = CIRCUITS[name]()
    n, ops = O.from_circuit(circuit, q)
    want, _ = O.measurement_distribution(n, ops)
    tp = trajectories.TrajectoryProgram(circuit, q)
    got = _engine_distribution(tp)
    assert abs(sum(got.values()) - 1.0) < 1e-12
    for bits in set(want) | set(got):
        assert abs(want.get(bits, 0.0) - got.get(bits, 0.0)) < 1e-12, (bits, want.get(bits), got.get(bits))


@pytest.mark.parametrize("name", sorted(CIRCUITS))
def test_simulate_returns_the_collapsed_state_of_its_outcomes(backend, name):
    circuit, q = CIRCUITS[name]()
    n, ops = O.from_circuit(circuit, q)
    _, finals = O.measurement_distribution(n, ops)
    sim = cirq.DensityMatrixSimulator(seed=11)
    spec = trajectories.TrajectoryProgram(circuit, q).measurements
    for _ in range(3):
        res = sim.simulate(circuit, qubit_order=q)
        masked = tuple(int(b) for key, _, _ in spec for b in res.measurements[key])
        rho = finals[_raw_bits(spec, masked)]
        assert np.abs(res.final_density_matrix - rho).max() < 1e-12


@pytest.mark.parametrize("name", sorted(CIRCUITS))
def test_run_samples_the_oracle_distribution(backend, name):
    circuit, q = CIRCUITS[name]()
    n, ops = O.from_circuit(circuit, q)
    want, _ = O.measurement_distribution(n, ops)
    reps = 20000
    res = cirq.DensityMatrixSimulator(seed=5).run(circuit, repetitions=reps)
    spec = trajectories.TrajectoryProgram(circuit, q).measurements
    cols = np.concatenate([res.measurements[key] for key, _, _ in spec], axis=1)
    assert cols.shape[0] == reps
    uniq, counts = np.unique(cols, axis=0, return_counts=True)
    seen = {_raw_bits(spec, tuple(int(b) for b in row)): c for row, c in zip(uniq, counts)}
    assert set(seen) <= {k for k, v in want.items() if v > 0}
    for bits, p in want.items():
        sigma = np.sqrt(max(p * (1 - p), 1e-12) / reps)
        assert abs(seen.get(bits, 0) / reps - p) < 5 * sigma + 2.5 / reps, (bits, p, seen.get(bits, 0) / reps)   # + Poisson tail of rare strings


def test_oracle_run_repeat_agrees_with_its_own_tree():
    """The restated per-repetition loop and the branch walk are two readings of the same cirq code: pin one on the other."""
    circuit, q = circuit_reused_qubit()
    n, ops = O.from_circuit(circuit, q)
    want, _ = O.measurement_distribution(n, ops)
    reps = 3000
    rec = O.run_repeat(n, ops, reps, np.random.RandomState(3))
    cols = np.concatenate([rec[k] for k in ("a", "b", "c")], axis=1)
    uniq, counts = np.unique(cols, axis=0, return_counts=True)
    for row, c in zip(uniq, counts):
        p = want[tuple(int(b) for b in row)]
        assert abs(c / reps - p) < 5 * np.sqrt(p * (1 - p) / reps) + 1e-9


def test_terminal_only_circuits_keep_the_evolve_once_path(backend):
    q = cirq.LineQubit.range(2)
    circuit = cirq.Circuit([cirq.H(q[0]), cirq.CNOT(q[0], q[1]), cirq.measure(q[0], q[1], key="z")])
    assert circuit.are_all_measurements_terminal()
    res = cirq.DensityMatrixSimulator(seed=1).run(circuit, repetitions=400)
    assert set(res.histogram(key="z")) <= {0, 3}
    assert not trajectories._CACHE        # the trajectory machinery was not involved


def test_repeated_keys_and_dephasing(backend):
    q = cirq.LineQubit.range(2)
    ops = [cirq.H(q[0]), cirq.CNOT(q[0], q[1]), cirq.measure(q[0], key="M"), cirq.H(q[0]), cirq.measure(q[0], key="M")]
    with pytest.raises(ValueError, match="Measurement key M repeated"):
        cirq.DensityMatrixSimulator().run(cirq.Circuit(ops), repetitions=2)
    # ignore_measurement_results: the measurement acts as the dephasing channel {|0><0|, |1><1|}, nothing is drawn
    ops[-1] = cirq.measure(q[0], key="N")
    circuit = cirq.Circuit(ops)
    rho = cirq.DensityMatrixSimulator(ignore_measurement_results=True).simulate(circuit, qubit_order=q).final_density_matrix
    p0, p1 = np.diag([1.0, 0.0]).astype(complex), np.diag([0.0, 1.0]).astype(complex)
    n, oops = O.from_circuit(circuit, q)
    oops = [("k", qs, [p0, p1]) if kind == "m" else (kind, qs, payload) for kind, qs, payload in oops]
    assert np.abs(rho - O.simulate(n, oops)).max() < 1e-12


def test_measurement_before_any_gate_and_structure_cache(backend):
    q = cirq.LineQubit.range(2)
    def build(theta):
        return cirq.Circuit([cirq.measure(q[1], key="first"), cirq.ry(theta)(q[0]), cirq.CNOT(q[0], q[1]), cirq.measure(q[0], key="mid"),
                             cirq.H(q[1]), cirq.measure(q[1], key="last")])
    sim = cirq.DensityMatrixSimulator(seed=2)
    for theta in (0.3, 1.2):
        res = sim.run(build(theta), repetitions=4000)
        assert not res.measurements["first"].any()
        assert abs(res.measurements["mid"].mean() - np.sin(theta / 2) ** 2) < 0.04
    assert len(trajectories._CACHE) == 1       # two angles, one compiled set of programs


def test_qpsamp_measurement_dialect_matches_the_projector_matrices(backend):
    """basis_ops_sim (MeasurementGate + post-selection, QP_QEM_lib.py:191-243) against basis_ops (Pz matrices, :246-304) on
    a 3-qubit SWAP test with projector operations drawn on purpose: exact through the outcome tree, statistical through run()."""
    from vqe_errormitigation_b200 import qp_qem_lib as Q
    qubits = cirq.LineQubit.range(3)
    samp = Q.QPSamp(Q.swap_circuit(1, qubits), "depolarize", 1e-2)
    n_gates = sum(1 for op in samp.circuit.all_operations() if not isinstance(op.gate, cirq.MeasurementGate))
    row = np.zeros(n_gates, np.uint8)
    one_q = [s for s, op in enumerate(o for o in samp.circuit.all_operations() if not isinstance(o.gate, cirq.MeasurementGate))
             if len(op.qubits) == 1]
    two_q = [s for s in range(n_gates) if s not in one_q]
    row[one_q[0]] = 12          # pi_z
    row[one_q[1]] = 10          # pi_x: rz rx M rx3 rz3
    row[two_q[0]] = 16 * 11 + 4  # pi_y on the first qubit, R_x on the second
    circuit, keys = samp.sampled_circuit_sim(row)
    assert len(keys) == 3 and not circuit.are_all_measurements_terminal()
    # exact: unnormalised P(all inserted read 0, final = b) from the outcome tree == the Pz-matrix variant state's diagonal
    # summed over the other qubits (the terminal measurement may or may not be the last operation of the moment order)
    tp = trajectories.TrajectoryProgram(circuit, qubits)
    spec_keys = [m[0] for m in tp.measurements]
    got = _engine_distribution(tp)
    mat = samp._variant_program().probabilities(None, row[None, :], renormalise=False)[0]
    for b in (0, 1):
        bits = tuple(b if k == "M" else 0 for k in spec_keys)
        assert abs(got[bits] - sum(mat[x] for x in range(8) if ((x >> 2) & 1) == b)) < 1e-12
    # statistical: the post-selected estimator
    want = sum(mat[x] * (1 - 2 * ((x >> 2) & 1)) for x in range(8))
    reps = 40000
    got = samp.postselected_expectation(row, reps, cirq.DensityMatrixSimulator(seed=9))
    assert abs(got - want) < 5 * np.sqrt(mat.sum() / reps)


def test_wide_mid_circuit_measurement_is_split_into_groups(backend):
    """Seven qubits measured at once in the middle of a circuit: more bits than one grouped batch evaluates
    (trajectories.MAX_GROUP_BITS), so the sites are walked as two levels - the joint distribution must not notice."""
    q = cirq.LineQubit.range(7)
    ops = [cirq.ry(0.4 + 0.3 * i)(q[i]) for i in range(7)] + [cirq.CNOT(q[i], q[i + 1]) for i in range(0, 6, 2)]
    ops += [cirq.measure(*q, key="all7"), cirq.H(q[0]), cirq.CNOT(q[0], q[6]), cirq.measure(q[0], q[6], key="end")]
    circuit = cirq.Circuit(_noise(ops, 0.01))
    tp = trajectories.TrajectoryProgram(circuit, q)
    assert [len(g.sites) for g in tp.groups] == [trajectories.MAX_GROUP_BITS, 1] and len(tp.final_sites) == 2
    n, oops = O.from_circuit(circuit, q)
    want, _ = O.measurement_distribution(n, oops)
    got = _engine_distribution(tp)
    assert abs(sum(got.values()) - 1.0) < 1e-12
    worst = max(abs(want.get(b, 0.0) - got.get(b, 0.0)) for b in set(want) | set(got))
    assert worst < 1e-12, worst


def test_simulate_with_unmeasured_qubits_at_the_final_measurement(monkeypatch):
    """Found by tools/fuzz_lowering.py --trajectories: when the circuit ends in a measurement of SOME qubits, the collapsed
    state of ``simulate`` is divided by the marginal probability of the measured bits (cirq: ``out /= probs[result]`` per
    measurement gate), not by the probability of the whole basis state the final draw landed on.  CPU double only - the GPU
    runs the same host code."""
    oracle_backend.install(monkeypatch)
    trajectories._CACHE.clear()
    engine._PROGRAM_CACHE.clear()
    q = cirq.LineQubit.range(4)
    ops = [cirq.H(q[1]), cirq.measure(q[2], q[3], q[0], key="m0"), cirq.depolarize(0.03).on(q[2]), cirq.H(q[3]),
           cirq.CNOT(q[0], q[2]), cirq.ry(0.4)(q[0]), cirq.measure(q[1], q[0], key="m1")]
    circuit = cirq.Circuit(ops)
    n, oracle_ops = O.from_circuit(circuit, q)
    _, finals = O.measurement_distribution(n, oracle_ops)
    spec = trajectories.TrajectoryProgram(circuit, q).measurements
    sim = cirq.DensityMatrixSimulator(seed=3)
    seen = set()
    for _ in range(12):
        res = sim.simulate(circuit, qubit_order=q)
        masked = tuple(int(b) for key, _, _ in spec for b in res.measurements[key])
        seen.add(masked)
        rho = finals[_raw_bits(spec, masked)]
        assert abs(np.trace(res.final_density_matrix) - 1.0) < 1e-12
        assert np.abs(res.final_density_matrix - rho).max() < 1e-12
    assert len(seen) >= 2
    trajectories._CACHE.clear()
    engine._PROGRAM_CACHE.clear()
