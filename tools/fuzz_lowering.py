"""Seed sweep of the random lowering checks of tests/test_lowering.py (CPU only: the dumped plan tables are replayed by
tests/plan_interp.py and compared with the oracle).  ``python tools/fuzz_lowering.py [first_seed] [n_seeds] [--stream | --ansatz | --trajectories | --circuits | --mitigation | --tile] [--opt key=value] [--qubits N]`` prints one line
per failure and a summary; exit code 1 when anything failed."""
import os
import sys
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import plan_interp as PI                                                      # noqa: E402
from oracle import dm_oracle as O                                             # noqa: E402
from test_gpu_parity import random_circuit, random_hamiltonian, resolve       # noqa: E402
from test_lowering import pauli_vector                                        # noqa: E402
from workloads import build_program                                           # noqa: E402

TOL = 1e-10
STREAM_QUBITS = 8          # --qubits N for --stream (8 ... 10; the interpreter takes ~6 s per 9-qubit seed)
PLAN_OPTIONS = {}          # --opt key=value (plan options of dmx_plan_set_option, e.g. fuse=0) for the default and --ansatz modes
_build_program = build_program


def build_program(ops, n, hamiltonian=None, **options):          # noqa: F811 - the sweep's own wrapper
    merged = dict(PLAN_OPTIONS)
    merged.update(options)
    return _build_program(ops, n, hamiltonian, **merged)
CLIFF1 = [O.H, O.rx(np.pi / 2), O.rx(-np.pi / 2), O.rx(np.pi), O.rz(np.pi / 2), O.X, O.Z]


def clifford_rotation_circuit(rng, n, length, n_params):
    ops = []
    for _ in range(length):
        kind = rng.integers(0, 5 if n_params > 1 else 4)
        if kind == 0:
            ops.append(("u", (int(rng.integers(n)),), CLIFF1[int(rng.integers(len(CLIFF1)))]))
        elif kind == 1 and n > 1:
            a, b = rng.choice(n, 2, replace=False)
            ops.append(("u", (int(a), int(b)), O.CNOT))
            ops.append(("k", (int(a), int(b)), O.two_qubit_depolarize(0.01)))
        elif kind == 2:
            ops.append(("k", (int(rng.integers(n)),), O.asymmetric_depolarize(0.01, 0.02, 0.005)))
        elif kind == 3:
            name = f"t{int(rng.integers(n_params))}"
            ops.append(("r", (int(rng.integers(n)),), (int(rng.integers(3)), name, float(rng.choice([1.0, -1.0, 2.0])), float(rng.choice([0.0, 0.0, 0.3])))))
        else:
            ops.append(("u", (int(rng.integers(n)),), O.rz(float(rng.normal()))))
    return ops


def one_seed(seed):
    rng = np.random.default_rng(seed)
    checked = []
    # (1) general circuits: fused blocks and packed device ops
    n = int(rng.integers(1, 6))
    ops = random_circuit(rng, n, depth=int(rng.integers(5, 40)))
    terms = random_hamiltonian(rng, n, int(rng.integers(1, 12)))
    ham = O.pauli_terms_to_matrix(n, *terms)
    prog = build_program(ops, n, terms)
    plan = PI.parse(prog.dump())
    vals = rng.normal(size=max(prog.n_params, 1))
    rho = O.simulate(n, resolve(ops, prog.param_names, vals))
    want_r = pauli_vector(rho, n)
    for name, r in (("blocks", PI.run_blocks(plan, vals)), ("devops", PI.run_devops(plan, vals))):
        assert np.max(np.abs(r - want_r)) < TOL, (name, n)
        assert abs(PI.energy(plan, r) - O.energy_sparse(rho, ham)) < TOL, (name, "energy")
    checked.append("general")
    # (2) Clifford + rotation circuits: frame tables, compact frame, thread-per-circuit tables
    for n_params in (1, 3):
        n = int(rng.integers(2, 5))
        ops = clifford_rotation_circuit(rng, n, int(rng.integers(8, 45)), n_params)
        terms = random_hamiltonian(rng, n, int(rng.integers(1, 12)))
        prog = build_program(ops, n, terms)
        plan = PI.parse(prog.dump())
        vals = rng.normal(size=max(prog.n_params, 1))
        want = O.energy_sparse(O.simulate(n, resolve(ops, prog.param_names, vals)), O.pauli_terms_to_matrix(n, *terms))
        if prog.info["path"] == 1 and "fsegs" in plan and "f_init" in plan:
            assert abs(PI.run_frame(plan, vals) - want) < TOL, ("frame", n, n_params)
            checked.append("frame")
        if "f32" in plan:
            assert abs(PI.run_frame32(plan, vals) - want) < TOL, ("frame32", n, n_params)
            checked.append("f32")
        if "f16" in plan and prog.n_params == 1:
            assert abs(PI.run_frame16t(plan, vals) - want) < TOL, ("frame16t", n)
            checked.append("f16")
            if plan["f16"]["scaled"]:
                assert abs(PI.run_frame16t_scaled(plan, vals) - want) < TOL, ("frame16t scaled", n)
                checked.append("f16s")
    return checked


def pauli_exponential_gates(rng, n, n_strings, alpha=None):
    """UCCSD-shaped gate list (i_wave_function.py:158-189): HF-like rx(pi) flips, then per random Pauli string the basis change,
    the CNOT ladder over its support, rz(2 sign alpha) on the last qubit, and everything undone.  alpha=None keeps the symbol."""
    gates = [("u", (q,), O.rx(np.pi)) for q in range(n) if rng.random() < 0.5]
    for _ in range(n_strings):
        support = sorted(int(q) for q in rng.choice(n, size=int(rng.integers(1, n + 1)), replace=False))
        letters = [str(rng.choice(["X", "Y", "Z"])) for _ in support]
        sign = float(rng.choice([1.0, -1.0]))
        for q, p in zip(support, letters):
            if p != "Z":
                gates.append(("u", (q,), O.H if p == "X" else O.rx(-np.pi / 2)))
        for a, b in zip(support[:-1], support[1:]):
            gates.append(("u", (a, b), O.CNOT))
        if alpha is None:
            gates.append(("r", (support[-1],), (2, "alpha0", 2.0 * sign, 0.0)))
        else:
            gates.append(("u", (support[-1],), O.rz(2.0 * sign * alpha)))
        for a, b in reversed(list(zip(support[:-1], support[1:]))):
            gates.append(("u", (a, b), O.CNOT))
        for q, p in zip(support, letters):
            if p != "Z":
                gates.append(("u", (q,), O.H if p == "X" else O.rx(np.pi / 2)))
    return gates


def pauli_noise(rng):
    n1 = [[O.depolarize(1e-3)], [O.asymmetric_depolarize(1e-3, 2e-3, 5e-4)], [O.bit_flip(1e-3), O.depolarize(1e-3)], [O.phase_flip(2e-3)]][int(rng.integers(4))]
    n2 = [[O.two_qubit_depolarize(1e-3)], [O.two_qubit_pauli_channel(1e-3, 1e-3, 6e-3)], []][int(rng.integers(3))]
    return n1, n2


def one_ansatz_seed(seed):
    """Thread-per-circuit tables (dmx_frame16t_kernel / dmx_frame16v_kernel) on UCCSD-shaped circuits with Pauli noise."""
    rng = np.random.default_rng(seed)
    checked = []
    n = int(rng.integers(2, 5))
    n_strings = int(rng.integers(1, 9))
    n1, n2 = pauli_noise(rng)
    terms = random_hamiltonian(rng, n, int(rng.integers(1, 16)))
    ham = O.pauli_terms_to_matrix(n, *terms)
    # (a) one symbol: parameter sweep
    gates = pauli_exponential_gates(np.random.default_rng(seed + 10 ** 6), n, n_strings)
    ops = O.with_noise(gates, n1, n2)
    prog = build_program(ops, n, terms)
    plan = PI.parse(prog.dump())
    for val in rng.normal(size=2):
        want = O.energy_sparse(O.simulate(n, resolve(ops, prog.param_names, [val])), ham)
        assert abs(PI.energy(plan, PI.run_devops(plan, [val])) - want) < TOL, ("devops", n)
        if prog.info["path"] == 1 and "fsegs" in plan and "f_init" in plan:
            assert abs(PI.run_frame(plan, [val]) - want) < TOL, ("frame", n)
        if "f32" in plan:
            assert abs(PI.run_frame32(plan, [val]) - want) < TOL, ("frame32", n)
        if "f16" in plan and not plan["f16"]["const"]:
            assert abs(PI.run_frame16t(plan, [val]) - want) < TOL, ("frame16t", n)
            checked.append("f16t")
            if plan["f16"]["scaled"]:
                assert abs(PI.run_frame16t_scaled(plan, [val]) - want) < TOL, ("frame16t scaled", n)
                checked.append("f16ts")
    # (b) the same circuit resolved, a variant slot after every gate: QP identifier rows
    alpha = float(rng.normal() * 0.3)
    gates = pauli_exponential_gates(np.random.default_rng(seed + 10 ** 6), n, n_strings, alpha=alpha)
    basis = O.basis_operations()
    vops = []
    for slot, (kind, qs, pl) in enumerate(gates):
        noise = n1 if len(qs) == 1 else n2
        vops.append((kind, qs, pl))
        vops += [("k", qs, kr) for kr in noise]
        vops.append(("v", qs, (basis, slot)))
        vops += [("ck", qs, (kr, slot)) for kr in noise]
    prog = build_program(vops, n, terms)
    plan = PI.parse(prog.dump())
    arity = [len(g[1]) for g in gates]
    for trial in range(6):
        v = np.zeros(len(gates), np.uint8)
        for slot in rng.choice(len(gates), size=min(len(gates), int(rng.integers(0, 9))), replace=False):
            if trial < 4:          # Pauli corrections (what the reference distribution draws)
                v[slot] = int(rng.integers(1, 4)) if arity[slot] == 1 else 16 * int(rng.integers(0, 4)) + int(rng.integers(0, 4))
            else:                  # any basis operation
                v[slot] = int(rng.integers(0, 16)) if arity[slot] == 1 else int(rng.integers(0, 256))
        want = O.energy_sparse(O.simulate(n, O.variant_ops(gates, n1, n2, v)), ham)
        scale = max(1.0, abs(want))
        assert abs(PI.energy(plan, PI.run_devops(plan, None, v)) - want) < TOL * scale, ("variant devops", n)
        checked.append("var")
        for name, fn in (("segments", PI.run_segments), ("frame", PI.run_frame), ("frame32", PI.run_frame32)):
            if name == "frame32" and "f32" not in plan:
                continue
            if name != "segments" and ("fsegs" not in plan or "f_init" not in plan):
                continue
            if name == "segments" and "e_w" not in plan:          # plan option segments=0: no segment tables were built
                continue
            got = fn(plan, None, v)
            if got is not None:
                assert abs(got - want) < TOL * scale, ("variant " + name, n)
        got = PI.run_frame16v(plan, v) if "f32" in plan else None
        if got is not None:
            assert abs(got - want) < TOL * scale, ("frame16v", n)
            checked.append("f16v")
    return checked


def one_trajectory_seed(seed):
    """Mid-circuit measurement (vqe_errormitigation_b200/trajectories.py) on the CPU test double: the outcome tree of a random
    circuit with measurements anywhere against the oracle's explicit projections (tests/test_trajectories.py's check)."""
    import pytest
    import oracle_backend
    import test_trajectories as TT
    from vqe_errormitigation_b200 import cirq_compat as cirq
    from vqe_errormitigation_b200 import engine, trajectories
    from vqe_errormitigation_b200.error_mitigation import IBasisGateSingle
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 5))
    q = cirq.LineQubit.range(n)
    ops, n_meas = [], 0
    for _ in range(int(rng.integers(3, 16))):
        kind = int(rng.integers(0, 7))
        a = int(rng.integers(n))
        if kind == 0:
            ops.append([cirq.H, cirq.X, cirq.Z][int(rng.integers(3))](q[a]))
        elif kind == 1:
            ops.append([cirq.rx, cirq.ry, cirq.rz][int(rng.integers(3))](float(rng.normal()))(q[a]))
        elif kind == 2 and n > 1:
            b = int(rng.choice([x for x in range(n) if x != a]))
            ops.append(cirq.CNOT(q[a], q[b]))
        elif kind == 3:
            ops.append(cirq.depolarize(0.03).on(q[a]) if rng.random() < 0.7 else cirq.amplitude_damp(0.1).on(q[a]))
        elif kind == 4 and rng.random() < 0.3:
            ops.append(IBasisGateSingle(np.array([[1.0, 0.0], [0.0, float(rng.uniform(0.3, 0.9))]], complex)).on(q[a]))
        elif n_meas < 4:
            k = int(rng.integers(1, min(n, 3) + 1))
            qs = [q[int(x)] for x in rng.choice(n, size=k, replace=False)]
            invert = tuple(bool(rng.integers(2)) for _ in qs) if rng.random() < 0.4 else ()
            ops.append(cirq.measure(*qs, key=f"m{n_meas}", invert_mask=invert))
            n_meas += 1
    if n_meas == 0:
        ops.insert(len(ops) // 2, cirq.measure(q[0], key="m0"))
    circuit = cirq.Circuit(ops)
    mp = pytest.MonkeyPatch()
    try:
        oracle_backend.install(mp)
        trajectories._CACHE.clear()
        engine._PROGRAM_CACHE.clear()
        nq, oops = O.from_circuit(circuit, q)
        want, finals = O.measurement_distribution(nq, oops)
        tp = trajectories.TrajectoryProgram(circuit, q)
        got = TT._engine_distribution(tp)
        assert abs(sum(got.values()) - 1.0) < 1e-10, "distribution does not sum to 1"
        for bits in set(want) | set(got):
            assert abs(want.get(bits, 0.0) - got.get(bits, 0.0)) < 1e-10, ("outcome probability", bits)
        res = cirq.DensityMatrixSimulator(seed=seed).simulate(circuit, qubit_order=q)
        masked = tuple(int(b) for key, _, _ in tp.measurements for b in res.measurements[key])
        rho = finals[TT._raw_bits(tp.measurements, masked)]
        assert np.abs(res.final_density_matrix - rho).max() < 1e-10, "collapsed final state"
        # run() records on a classical circuit (X / CNOT / SWAP only): every shot must read the bits a classical simulation
        # gives, invert masks applied, whether the measurements are terminal (evolve once + sample) or not (trajectories)
        m = int(rng.integers(1, 5))
        cq = cirq.LineQubit.range(m)
        bits = [0] * m
        cops, expect, n_keys = [], {}, 0
        terminal_only = bool(rng.integers(2))
        body = int(rng.integers(2, 14))
        for step in range(body):
            a = int(rng.integers(m))
            others = [x for x in range(m) if x != a]
            kind = int(rng.integers(0, 4))
            if kind == 0 or not others:
                cops.append(cirq.X(cq[a]))
                bits[a] ^= 1
            elif kind == 1:
                b = int(rng.choice(others))
                cops.append(cirq.CNOT(cq[a], cq[b]))
                bits[b] ^= bits[a]
            elif kind == 2:
                b = int(rng.choice(others))
                cops.append(cirq.SWAP(cq[a], cq[b]))
                bits[a], bits[b] = bits[b], bits[a]
            elif not terminal_only and n_keys < 3:
                qs = [int(x) for x in rng.choice(m, size=int(rng.integers(1, m + 1)), replace=False)]
                invert = tuple(bool(rng.integers(2)) for _ in qs) if rng.random() < 0.5 else ()
                cops.append(cirq.measure(*[cq[x] for x in qs], key=f"k{n_keys}", invert_mask=invert))
                expect[f"k{n_keys}"] = [bits[x] ^ (1 if i < len(invert) and invert[i] else 0) for i, x in enumerate(qs)]
                n_keys += 1
        for _ in range(int(rng.integers(1, 3))):
            qs = [int(x) for x in rng.choice(m, size=int(rng.integers(1, m + 1)), replace=False)]
            invert = tuple(bool(rng.integers(2)) for _ in qs) if rng.random() < 0.5 else ()
            cops.append(cirq.measure(*[cq[x] for x in qs], key=f"k{n_keys}", invert_mask=invert))
            expect[f"k{n_keys}"] = [bits[x] ^ (1 if i < len(invert) and invert[i] else 0) for i, x in enumerate(qs)]
            n_keys += 1
        trajectories._CACHE.clear()
        engine._PROGRAM_CACHE.clear()
        result = cirq.DensityMatrixSimulator(seed=seed).run(cirq.Circuit(cops), repetitions=3)
        for key, want_bits in expect.items():
            got_bits = np.asarray(result.measurements[key])
            assert got_bits.shape == (3, len(want_bits)) and (got_bits == np.array(want_bits)).all(), ("run() record", key)
    finally:
        mp.undo()
        trajectories._CACHE.clear()
        engine._PROGRAM_CACHE.clear()
    return ["tree", "simulate", "records"]


def one_circuit_seed(seed):
    """The circuit-facing side (engine.lower_circuit behind DensityMatrixSimulator / QPU-style energy programs) on the CPU test
    double: random circuits over the whole gate vocabulary of cirq_compat, symbols included, against the oracle's own reading of
    the same circuit (O.from_circuit)."""
    import pytest
    import sympy
    import oracle_backend
    from vqe_errormitigation_b200 import cirq_compat as cirq
    from vqe_errormitigation_b200 import engine, trajectories
    from vqe_errormitigation_b200.error_mitigation import (IBasisGateSingle, IBasisGateTwo, SingleQubitPauliChannel, TwoQubitDepolarizingChannel,
                                                           TwoQubitPauliChannel)
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 6))
    kind_q = int(rng.integers(3))
    q = cirq.LineQubit.range(n) if kind_q == 0 else [cirq.GridQubit(i // 2, i % 2) for i in range(n)] if kind_q == 1 else \
        [cirq.NamedQubit(f"q{chr(97 + i)}") for i in range(n)]
    symbols = [sympy.Symbol(f"s{i}") for i in range(3)]
    values = {f"s{i}": float(rng.normal()) for i in range(3)}
    ops = []
    for _ in range(int(rng.integers(3, 30))):
        kind = int(rng.integers(0, 12))
        a = int(rng.integers(n))
        others = [x for x in range(n) if x != a]
        b = int(rng.choice(others)) if others else None
        if kind == 0:
            ops.append([cirq.X, cirq.Y, cirq.Z, cirq.H, cirq.S, cirq.T, cirq.I][int(rng.integers(7))](q[a]))
        elif kind == 1:
            ops.append([cirq.rx, cirq.ry, cirq.rz][int(rng.integers(3))](float(rng.normal()))(q[a]))
        elif kind == 2:
            angle = float(rng.choice([1.0, -2.0, 0.5])) * symbols[int(rng.integers(3))] + float(rng.choice([0.0, 0.3]))
            ops.append([cirq.rx, cirq.ry, cirq.rz][int(rng.integers(3))](angle)(q[a]))
        elif kind == 3:
            ops.append(([cirq.X, cirq.Y, cirq.Z, cirq.H][int(rng.integers(4))] ** float(rng.uniform(-1, 1)))(q[a]))
        elif kind == 4 and b is not None:
            ops.append([cirq.CNOT, cirq.CZ, cirq.SWAP][int(rng.integers(3))](q[a], q[b]))
        elif kind == 5 and b is not None:
            ops.append(([cirq.CNOT, cirq.CZ, cirq.SWAP][int(rng.integers(3))] ** float(rng.uniform(-1, 1)))(q[a], q[b]))
        elif kind == 6 and len(others) >= 2:
            b2, c2 = (int(x) for x in rng.choice(others, 2, replace=False))
            gate = [cirq.CSWAP, cirq.TOFFOLI][int(rng.integers(2))](q[a], q[b2], q[c2])
            ops.append(gate._decompose_() if hasattr(gate, "_decompose_") and rng.random() < 0.5 else gate)
        elif kind == 7:
            ops.append([cirq.depolarize(0.02), cirq.bit_flip(0.03), cirq.phase_flip(0.04), cirq.amplitude_damp(0.1), cirq.phase_damp(0.2),
                        cirq.asymmetric_depolarize(0.01, 0.02, 0.03), SingleQubitPauliChannel(0.01, 0.0, 0.05)][int(rng.integers(7))](q[a]))
        elif kind == 8 and b is not None:
            ops.append([TwoQubitDepolarizingChannel(0.03), TwoQubitPauliChannel(0.01, 0.02, 0.03)][int(rng.integers(2))](q[a], q[b]))
        elif kind == 9:
            m = rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2))
            ops.append(IBasisGateSingle(m / np.linalg.norm(m, 2)).on(q[a]))
        elif kind == 10 and b is not None:
            m = np.linalg.qr(rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4)))[0]
            ops.append((IBasisGateTwo(m) if rng.random() < 0.5 else cirq.MatrixGate(m)).on(q[a], q[b]))
        elif kind == 11 and rng.random() < 0.3:
            ops.append(cirq.MatrixGate(np.linalg.qr(rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2)))[0]).on(q[a]))
    if not ops:
        ops.append(cirq.H(q[0]))
    strategy = [cirq.InsertStrategy.EARLIEST, cirq.InsertStrategy.NEW, cirq.InsertStrategy.INLINE][int(rng.integers(3))]
    circuit = cirq.Circuit(ops, strategy=strategy)
    order = [q[i] for i in rng.permutation(n)] if rng.random() < 0.3 else None
    used = list(order) if order is not None else sorted(circuit.all_qubits())
    resolver = cirq.ParamResolver(values)
    resolved = cirq.resolve_parameters(circuit, resolver)
    nq, oracle_ops = O.from_circuit(resolved, used)
    want = O.simulate(nq, oracle_ops)
    scale = max(1.0, float(np.abs(want).max()))
    mp = pytest.MonkeyPatch()
    try:
        oracle_backend.install(mp)
        trajectories._CACHE.clear()
        engine._PROGRAM_CACHE.clear()
        got = cirq.DensityMatrixSimulator().simulate(circuit, param_resolver=resolver, qubit_order=order).final_density_matrix
        assert got.shape == want.shape and np.abs(got - want).max() < 1e-10 * scale, "simulate (resolver)"
        again = cirq.DensityMatrixSimulator().simulate(resolved, qubit_order=order).final_density_matrix
        assert np.abs(again - want).max() < 1e-10 * scale, "simulate (resolved circuit)"
        # the symbolic circuit as one plan with parameter columns, and its energy on random Pauli terms
        terms = random_hamiltonian(rng, nq, int(rng.integers(1, 8)))
        prog, _, _ = engine.compile_circuit(circuit, used, hamiltonian=terms)
        row = None if not prog.n_params else np.array([[values[nm] for nm in prog.param_names]])
        e = float(prog.energies_host(row, batch=1)[0])
        e_want = float(np.real(np.trace(want @ O.pauli_terms_to_matrix(nq, *terms))))
        assert abs(e - e_want) < 1e-10 * scale, "energy (symbolic plan)"
    finally:
        mp.undo()
        trajectories._CACHE.clear()
        engine._PROGRAM_CACHE.clear()
    return ["simulate", "energy"]


def one_mitigation_seed(seed):
    """IErrorMitigationManager on the CPU test double: (1) the batched variant program against the reference-shaped
    ``get_updated_circuit`` of each row, simulated by the oracle; (2) on a one- or two-gate circuit the COMPLETE quasi-probability
    sum over all identifier rows reproduces the noiseless value (the estimator's defining identity)."""
    import itertools
    import pytest
    import oracle_backend
    from vqe_errormitigation_b200 import cirq_compat as cirq
    from vqe_errormitigation_b200 import engine
    from vqe_errormitigation_b200.data_containers.helper_interfaces.i_noise_wrapper import INoiseModel
    from vqe_errormitigation_b200.error_mitigation import (IErrorMitigationManager, SingleQubitPauliChannel, TwoQubitDepolarizingChannel,
                                                           TwoQubitPauliChannel)
    rng = np.random.default_rng(seed)

    def random_gate(q, n, allow_two=True):
        a = int(rng.integers(n))
        kind = int(rng.integers(0, 4 if (allow_two and n > 1) else 3))
        if kind == 0:
            return [cirq.H, cirq.X, cirq.S, cirq.T][int(rng.integers(4))](q[a])
        if kind in (1, 2):
            return [cirq.rx, cirq.ry, cirq.rz][int(rng.integers(3))](float(rng.normal()))(q[a])
        b = int(rng.choice([x for x in range(n) if x != a]))
        return [cirq.CNOT, cirq.CZ][int(rng.integers(2))](q[a], q[b])

    def random_model():
        p = float(rng.choice([1e-3, 1e-2, 5e-2]))
        n1 = [[cirq.depolarize(p)], [cirq.bit_flip(p), cirq.depolarize(p)], [SingleQubitPauliChannel(p, p, 6 * p)], [cirq.phase_flip(p)]][int(rng.integers(4))]
        n2 = [[TwoQubitDepolarizingChannel(p)], [TwoQubitPauliChannel(p, p, 6 * p)], []][int(rng.integers(3))]
        return INoiseModel(n1, n2, "fuzz")

    def objective_for(n):
        if rng.random() < 0.3:
            return None
        h = rng.normal(size=(1 << n, 1 << n))
        if rng.random() < 0.4:          # scipy sparse objective: `*` is the matrix product, the full Tr(H rho) (reference :382)
            import scipy.sparse
            g = h + 1j * rng.normal(size=h.shape)
            return scipy.sparse.csc_matrix(g + g.conj().T)
        return h + h.T

    def expected(circuit, objective, n, q):
        rho = O.simulate(n, O.from_circuit(circuit, q)[1])
        if objective is None:
            return O.energy_elementwise(rho, np.kron(np.eye(1 << (n - 1)), O.Z))
        if hasattr(objective, "toarray"):
            return float(np.real(np.trace(objective.toarray() @ rho)))
        return O.energy_elementwise(rho, objective)

    checked = []
    mp = pytest.MonkeyPatch()
    try:
        oracle_backend.install(mp)
        engine._PROGRAM_CACHE.clear()
        # (1) random rows of a longer circuit
        n = int(rng.integers(1, 4))
        q = cirq.LineQubit.range(n)
        clean = cirq.Circuit([random_gate(q, n) for _ in range(int(rng.integers(2, 10)))] + [cirq.H(x) for x in q])
        model, objective = random_model(), objective_for(n)
        mgr = IErrorMitigationManager(clean, model, objective)
        gate_ops = [op for op in clean.all_operations()]
        rows = np.zeros((6, len(gate_ops)), np.uint8)
        for r in range(1, 6):
            for s in rng.choice(len(gate_ops), size=min(len(gate_ops), int(rng.integers(1, 4))), replace=False):
                rows[r, s] = int(rng.integers(0, 16)) if len(gate_ops[s].qubits) == 1 else int(rng.integers(0, 256))
        values, _ = mgr.evaluate_variants(rows)
        for r in range(6):
            circuit = IErrorMitigationManager.get_updated_circuit(clean, model.get_operators, list(rows[r]))
            want = expected(circuit, objective, n, q)
            assert abs(values[r] - want) < 1e-10 * max(1.0, abs(want)), ("variant row", r)
        checked.append("rows")
        # (2) every identifier row of a tiny circuit (all 16 / 256 basis operations of every slot).  The complete
        # quasi-probability sum is NOT the ideal value exactly: the reference re-applies the noise after a non-identity
        # correction (error_mitigation.py:528-572), which leaves a residual of order p * (cost - 1) - bounded loosely below
        n = int(rng.integers(1, 3))
        q = cirq.LineQubit.range(n)
        if n == 1:
            ops = [random_gate(q, 1)] + ([random_gate(q, 1)] if rng.random() < 0.5 else [])
        else:
            ops = [[cirq.CNOT, cirq.CZ][int(rng.integers(2))](q[0], q[1])] if rng.random() < 0.5 else \
                [random_gate([q[0]], 1), random_gate([q[1]], 1)]
        clean = cirq.Circuit(ops)
        model, objective = random_model(), objective_for(n)
        mgr = IErrorMitigationManager(clean, model, objective)
        qps = mgr.quasi_probabilities()
        sizes = [len(qp) for qp in qps]
        rows = np.array(list(itertools.product(*[range(k) for k in sizes])), np.uint8)
        weights = np.prod([np.asarray(qp)[rows[:, s]] for s, qp in enumerate(qps)], axis=0)
        keep = np.ones(len(rows), bool)
        values, _ = mgr.evaluate_variants(rows)
        for r in range(len(rows)):
            want = expected(IErrorMitigationManager.get_updated_circuit(clean, model.get_operators, list(rows[r])), objective, n, q)
            assert abs(values[r] - want) < 1e-10 * max(1.0, abs(want)), ("tiny circuit row", list(rows[r]))
        ideal = expected(clean, objective, n, q)
        cost = float(np.prod([np.sum(np.abs(qp)) for qp in qps]))
        scale = 1.0 if objective is None else float(np.abs(objective.toarray() if hasattr(objective, "toarray") else np.diagonal(objective)).max())
        assert abs(float(np.dot(weights, values)) - ideal) < 3.0 * (cost - 1.0 + 1e-9) * 0.1 * max(scale, 1e-3) + 1e-9, "complete sum far from ideal"
        # and the signs / cost bookkeeping: cost * sign * |w| / cost == w
        signs = mgr.variant_signs(rows[keep])
        parity = np.prod([np.sign(np.asarray(qp)[rows[:, s]]) for s, qp in enumerate(qps)], axis=0)     # (the product itself can underflow)
        assert np.array_equal(signs, parity), "variant signs"
        assert abs(mgr.total_cost() - float(np.prod([np.sum(np.abs(qp)) for qp in qps]))) < 1e-12, "total cost"
        checked.append("sum")
    finally:
        mp.undo()
        engine._PROGRAM_CACHE.clear()
    return checked


def one_tile_seed(seed):
    """Whole-state tile plans (6 and 7 qubits, dmx_tile_kernel's packed device ops): energies on 30 random Pauli terms, with
    parameters and with a few variant slots."""
    rng = np.random.default_rng(seed)
    n = int(rng.integers(6, 8))
    ops = random_circuit(rng, n, depth=int(rng.integers(10, 40)))
    basis = O.basis_operations()
    slots = 0
    for _ in range(int(rng.integers(0, 4))):
        at = int(rng.integers(len(ops) + 1))
        ops.insert(at, ("v", (int(rng.integers(n)),), (basis, slots)))
        slots += 1
    terms = random_hamiltonian(rng, n, 30)
    ham = O.pauli_terms_to_matrix(n, *terms)
    prog = build_program(ops, n, terms)
    plan = PI.parse(prog.dump())
    vals = rng.normal(size=max(prog.n_params, 1))
    variant = np.array([int(rng.integers(0, 16)) for _ in range(slots)], np.uint8) if slots else None
    resolved = []
    for kind, qs, pl in resolve(ops, prog.param_names, vals):
        if kind == "v":
            resolved.append(("u", qs, pl[0][int(variant[pl[1]])]))
        else:
            resolved.append((kind, qs, pl))
    want = O.energy_sparse(O.simulate(n, resolved), ham)
    scale = max(1.0, abs(want))
    for name, fn in (("blocks", PI.run_blocks), ("devops", PI.run_devops)):
        got = PI.energy(plan, fn(plan, vals, variant))
        assert abs(got - want) < TOL * scale, (name, n)
    return ["tile%d" % n]


def one_stream_seed(seed):
    """The streaming (state in HBM) lowering on 8 qubits: pair / triple sweeps, wide monomial sweeps, both tile sizes, remapped
    or fixed tiles - numerics only (the sweep-count expectations of tests/test_lowering.py are about speed, not parity)."""
    import test_lowering as TL
    n = STREAM_QUBITS
    rng = np.random.default_rng(800 + seed)
    ops = TL._clifford_mix_circuit(rng, n, int(rng.integers(20, 80)))
    terms = random_hamiltonian(rng, n, int(rng.integers(1, 14)))
    ham = O.pauli_terms_to_matrix(n, *terms)
    checked = []
    want = vals = None
    for options in ({"tile_qubits": 6, "wide": 0}, {"tile_qubits": 6, "wide": 1}, {"tile_qubits": 7, "wide": 0}, {"tile_qubits": 7, "wide": 1},
                    {"tile_qubits": 7, "triples": 0}, {"tile_qubits": 6, "remap": 0}, {"tile_qubits": 5}):
        prog = build_program(ops, n, terms, **options)
        plan = PI.parse(prog.dump())
        assert prog.info["path"] == 2, ("not a streaming plan", options)
        if want is None:
            vals = rng.uniform(-1, 1, size=prog.n_params)
            want = O.energy_sparse(O.simulate(n, resolve(ops, prog.param_names, vals)), ham)
        assert abs(PI.energy(plan, PI.run_devops(plan, vals)) - want) < TOL, ("stream", options)
        checked.append("stream")
    return checked


def main():
    argv = list(sys.argv[1:])
    while "--opt" in argv:
        i = argv.index("--opt")
        key, value = argv[i + 1].split("=")
        PLAN_OPTIONS[key] = int(value)
        del argv[i:i + 2]
    if "--qubits" in argv:
        global STREAM_QUBITS
        i = argv.index("--qubits")
        STREAM_QUBITS = int(argv[i + 1])
        del argv[i:i + 2]
    args = [a for a in argv if not a.startswith("--")]
    first = int(args[0]) if len(args) > 0 else 1000
    count = int(args[1]) if len(args) > 1 else 100
    run = one_stream_seed if "--stream" in sys.argv else one_ansatz_seed if "--ansatz" in sys.argv else one_trajectory_seed if "--trajectories" in sys.argv else one_circuit_seed if "--circuits" in sys.argv else one_mitigation_seed if "--mitigation" in sys.argv else one_tile_seed if "--tile" in sys.argv else one_seed
    failures, tally = 0, {}
    for seed in range(first, first + count):
        try:
            for name in run(seed):
                tally[name] = tally.get(name, 0) + 1
        except Exception:                                                     # noqa: BLE001 - report and go on
            failures += 1
            print(f"seed {seed}: {traceback.format_exc().strip().splitlines()[-1]}", flush=True)
    print(f"seeds {first}..{first + count - 1}: {failures} failure(s); checks run {tally}")
    return 1 if failures else 0


if __name__ == "__main__":
    sys.exit(main())
