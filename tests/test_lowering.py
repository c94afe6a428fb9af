"""Host-side lowering (csrc/dmx_lower.cpp) verified on CPU: the dumped program — fused Pauli-transfer blocks,
the packed device-op tables and the segment-kernel tables — is interpreted with numpy (plan_interp.py) exactly
as the kernels are specified to execute it, and compared with the oracle."""
import json
import os

import numpy as np
import pytest

import plan_interp as PI
from oracle import dm_oracle as O
from test_gpu_parity import qp_noisy_variant_program, random_circuit, random_hamiltonian, random_unitary, resolve
from workloads import ALPHA_STAR, GOLDEN, build_program, h2_ops, h2_terms

TOL = 1e-11


def pauli_vector(rho, n):
    """r[P] = Tr(rho P) in the engine's index order (2 bits per qubit, code 0=I 1=Z 2=X 3=Y)."""
    paulis = [O.I2, O.Z, O.X, O.Y]
    out = np.zeros(4 ** n)
    for idx in range(4 ** n):
        m = np.eye(1)
        for q in range(n):
            m = np.kron(m, paulis[(idx >> (2 * (n - 1 - q))) & 3])
        out[idx] = np.trace(rho @ m).real
    return out


@pytest.mark.parametrize("noise", ["dep", "pauli", "none", "ampdamp"])
@pytest.mark.parametrize("options", [{}, {"fuse": 0}])
def test_h2_lowering(noise, options):
    n1, n2 = {"dep": ([O.depolarize(1e-3)], [O.two_qubit_depolarize(1e-3)]),
              "pauli": ([O.asymmetric_depolarize(1e-4, 1e-4, 6e-4)], [O.two_qubit_pauli_channel(1e-4, 1e-4, 6e-4)]),
              "none": ([], []), "ampdamp": ([O.amplitude_damp(1e-2)], [])}[noise]
    terms = h2_terms()
    ham = O.pauli_terms_to_matrix(4, *terms)
    prog = build_program(h2_ops(n1, n2), 4, terms, **options)
    plan = PI.parse(prog.dump())
    assert (plan["nseg"] == 8) == (noise != "ampdamp")
    if noise != "ampdamp" and not options:
        assert prog.info["n_blocks"] < 60      # 244 (or 122) ops fuse to ~49 gate+noise blocks
    for a in (0.0, ALPHA_STAR, -0.9):
        want = O.energy_sparse(O.simulate(4, h2_ops(n1, n2, alpha=a)), ham)
        assert abs(PI.energy(plan, PI.run_blocks(plan, [a])) - want) < TOL
        assert abs(PI.energy(plan, PI.run_devops(plan, [a])) - want) < TOL
        if plan["nseg"]:
            assert abs(PI.run_segments(plan, [a]) - want) < TOL
            assert abs(PI.run_frame(plan, [a]) - want) < TOL
            if noise != "ampdamp":
                assert plan["f32"]["n"] <= 24              # noisy / clean H2: at most 24 of 256 coefficients matter
            if "f32" in plan:
                assert abs(PI.run_frame32(plan, [a]) - want) < TOL
            if noise != "ampdamp":
                # thread-per-circuit form: the 8 UCCSD rotations touch 16 coefficients (rank-4 XOR pattern), 7 are constants
                assert plan["f16"]["n"] == 16 and all(0 < m < 16 for m in plan["f16"]["mask"])
                assert abs(PI.run_frame16t(plan, [a]) - want) < TOL
                assert plan["f16"]["scaled"] == 1 and abs(PI.run_frame16t_scaled(plan, [a]) - want) < TOL


def test_frame16t_tables_on_random_single_class_circuits():
    """One-parameter Clifford + Pauli-noise + rotation circuits: whenever the lowering offers the thread-per-circuit tables
    they reproduce the oracle; circuits whose rotations touch more than 16 coefficients must not offer them."""
    rng = np.random.default_rng(2026)
    cliff1 = [O.H, O.rx(np.pi / 2), O.rx(-np.pi / 2), O.rx(np.pi), O.rz(np.pi / 2), O.X, O.Z]
    offered = scaled = 0
    for trial in range(24):
        n = int(rng.integers(2, 5))
        ops = []
        for _ in range(int(rng.integers(10, 40))):
            kind = rng.integers(0, 4)
            if kind == 0:
                ops.append(("u", (int(rng.integers(n)),), cliff1[int(rng.integers(len(cliff1)))]))
            elif kind == 1:
                a, b = rng.choice(n, 2, replace=False)
                ops.append(("u", (int(a), int(b)), O.CNOT))
                ops.append(("k", (int(a), int(b)), O.two_qubit_depolarize(0.01)))
            elif kind == 2:
                ops.append(("k", (int(rng.integers(n)),), O.asymmetric_depolarize(0.01, 0.02, 0.005)))
            else:
                ops.append(("r", (int(rng.integers(n)),), (int(rng.integers(3)), "t0", 2.0, 0.0)))
        terms = random_hamiltonian(rng, n, 6)
        prog = build_program(ops, n, terms)
        plan = PI.parse(prog.dump())
        if "f16" not in plan:
            continue
        offered += 1
        assert plan["f16"]["n"] <= 16 and len(plan["fclasses"]) == 1
        scaled += plan["f16"]["scaled"]
        for val in rng.normal(size=3):
            want = O.energy_sparse(O.simulate(n, resolve(ops, prog.param_names, [val])), O.pauli_terms_to_matrix(n, *terms))
            assert abs(PI.run_frame16t(plan, [val]) - want) < TOL
            if plan["f16"]["scaled"]:
                assert abs(PI.run_frame16t_scaled(plan, [val]) - want) < TOL
    assert offered >= 3 and scaled >= 1


def test_frame_relabelling_keeps_h2_exchanges_inside_warps():
    """The 8 rotation axes of the H2 ansatz span a rank-4 GF(2) space, so all partner exchanges are warp shuffles."""
    prog = build_program(h2_ops([O.depolarize(1e-3)], [O.two_qubit_depolarize(1e-3)]), 4, h2_terms())
    plan = PI.parse(prog.dump())
    assert len(plan["fsegs"]) == 8 and all(f["shfl"] == 1 and 0 < f["mask"] < 32 for f in plan["fsegs"])
    assert plan["single_class"] == 1 and len(plan["fclasses"]) == 1


def test_frame_mixed_axes_random_clifford_rotation_circuits():
    """Random Clifford + Pauli-noise + rx/ry/rz circuits on 2-4 qubits: frame tables (shuffle and smem segments,
    several rotation classes) reproduce the oracle."""
    rng = np.random.default_rng(77)
    cliff1 = [O.H, O.rx(np.pi / 2), O.rx(-np.pi / 2), O.rx(np.pi), O.rz(np.pi / 2), O.X, O.Z]
    n_compact = 0
    for trial in range(6):
        n = int(rng.integers(2, 5))
        ops = []
        for _ in range(40):
            kind = rng.integers(0, 5)
            if kind == 0:
                ops.append(("u", (int(rng.integers(n)),), cliff1[int(rng.integers(len(cliff1)))]))
            elif kind == 1:
                a, b = rng.choice(n, 2, replace=False)
                ops.append(("u", (int(a), int(b)), O.CNOT))
                ops.append(("k", (int(a), int(b)), O.two_qubit_depolarize(0.01)))
            elif kind == 2:
                ops.append(("k", (int(rng.integers(n)),), O.asymmetric_depolarize(0.01, 0.02, 0.005)))
            elif kind == 3:
                ops.append(("r", (int(rng.integers(n)),), (int(rng.integers(3)), f"t{int(rng.integers(3))}", float(rng.choice([1.0, -1.0, 2.0])), float(rng.choice([0.0, 0.0, 0.3])))))
            else:
                ops.append(("u", (int(rng.integers(n)),), O.rz(float(rng.normal()))))
        terms = random_hamiltonian(rng, n, 10)
        prog = build_program(ops, n, terms)
        assert prog.info["path"] == 1, prog.dump().splitlines()[0]
        plan = PI.parse(prog.dump())
        vals = rng.normal(size=max(prog.n_params, 1))
        want = O.energy_sparse(O.simulate(n, resolve(ops, prog.param_names, vals)), O.pauli_terms_to_matrix(n, *terms))
        assert abs(PI.run_frame(plan, vals) - want) < TOL
        if "f32" in plan:
            n_compact += 1
            assert abs(PI.run_frame32(plan, vals) - want) < TOL
    assert n_compact >= 1


def test_canonical_cost_model_matches_survey():
    """74*28*256 + 74*6*256 + 48*6*256 flops for the noisy H2 circuit + energy epilogue."""
    prog = build_program(h2_ops([O.depolarize(1e-3)], [O.two_qubit_depolarize(1e-3)]), 4, h2_terms())
    assert prog.info["canonical_flops_per_circuit"] == 74 * 28 * 256 + 74 * 6 * 256 + 48 * 6 * 256 + (8 * 2 + 2 * 15) * 16


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5])
def test_random_circuit_lowering(n):
    rng = np.random.default_rng(40 + n)
    ops = random_circuit(rng, n, depth=30)
    terms = random_hamiltonian(rng, n, 9)
    ham = O.pauli_terms_to_matrix(n, *terms)
    prog = build_program(ops, n, terms)
    plan = PI.parse(prog.dump())
    vals = rng.normal(size=max(prog.n_params, 1))
    rho = O.simulate(n, resolve(ops, prog.param_names, vals))
    want_r = pauli_vector(rho, n)
    for r in (PI.run_blocks(plan, vals), PI.run_devops(plan, vals)):
        assert np.max(np.abs(r - want_r)) < TOL
        assert abs(PI.energy(plan, r) - O.energy_sparse(rho, ham)) < TOL


def test_variant_lowering_against_golden():
    with open(os.path.join(GOLDEN, "h2_noisy_energies.json")) as fh:
        gold = json.load(fh)
    n1, n2 = [O.bit_flip(1e-3), O.depolarize(1e-3)], [O.two_qubit_depolarize(1e-3)]
    prog, _ = qp_noisy_variant_program(n1, n2, h2_terms())
    plan = PI.parse(prog.dump())
    assert plan["nseg"] == 8 and prog.info["path"] == 1   # resolved rz gates close constant-coefficient segments
    general = 0
    for case in gold["variant_cases"]:
        v = np.array(case["identifier"])
        assert abs(PI.energy(plan, PI.run_blocks(plan, None, v)) - case["energy"]) < TOL
        assert abs(PI.energy(plan, PI.run_devops(plan, None, v)) - case["energy"]) < TOL
        seg = PI.run_segments(plan, None, v)
        frame = PI.run_frame(plan, None, v)
        if seg is None:
            general += 1        # non-diagonal correction op: kernel defers this row to the tile path
            assert frame is None
        else:
            assert abs(seg - case["energy"]) < TOL
            assert abs(frame - case["energy"]) < TOL
            assert abs(PI.run_frame32(plan, None, v) - case["energy"]) < TOL
            assert abs(PI.run_frame16v(plan, v) - case["energy"]) < TOL       # thread-per-row form (dmx_frame16v_kernel)
    assert general >= 1
    assert plan["f16"]["const"] == 1 and plan["f16"]["n"] == 16 and len(plan["f16"]["static"]) == 7
    # random rows of diagonal corrections (Pauli basis operations, indices 1-3 per qubit) against the oracle
    rng = np.random.default_rng(8)
    n_slots = prog.n_slots
    gates = O.h2_gate_list(ALPHA_STAR)
    for trial in range(12):
        v = np.zeros(n_slots, np.uint8)
        for s in rng.choice(n_slots, size=int(rng.integers(1, 6)), replace=False):
            nq = plan["slots"][int(s)]["nq"]
            v[s] = int(rng.integers(1, 4)) if nq == 1 else 16 * int(rng.integers(0, 4)) + int(rng.integers(0, 4))
        want = O.energy_sparse(O.simulate(4, O.variant_ops(gates, n1, n2, v)), O.pauli_terms_to_matrix(4, *h2_terms()))
        assert abs(PI.run_frame16v(plan, v) - want) < TOL, (trial, v.nonzero())


def test_streaming_schedule_partitions_all_blocks():
    n = 9
    ops = []
    for layer in range(2):
        for q in range(n):
            ops += [("r", (q,), (1, f"y{layer}_{q}", 1.0, 0.0)), ("k", (q,), O.depolarize(1e-3))]
        for q in range(n - 1):
            ops += [("u", (q, q + 1), O.CNOT), ("k", (q, q + 1), O.two_qubit_depolarize(1e-3))]
    rng = np.random.default_rng(1)
    terms = random_hamiltonian(rng, n, 5)
    n_pair_sweeps = None
    for tile, remap, triples in ((4, 1, 1), (5, 1, 1), (7, 0, 1), (6, 1, 0), (7, 1, 0), (6, 1, 1), (7, 1, 1)):
        prog = build_program(ops, n, terms, tile_qubits=tile, remap=remap, triples=triples)
        plan = PI.parse(prog.dump())
        assert prog.info["path"] == 2
        assert sum(p["nphases"] for p in plan["passes"]) == len(plan["phases"])
        sweeps = plan["sweeps3"] if (remap and triples and tile >= 6) else plan["sweeps"]
        assert sum(ph["nsweeps"] for ph in plan["phases"]) == len(sweeps)
        assert len(sweeps) < len(plan["blocks"])     # several fused blocks per shared-memory round trip
        if tile == 7 and remap:
            if not triples:
                n_pair_sweeps = len(sweeps)
            else:
                assert len(sweeps) < 0.75 * n_pair_sweeps        # a CNOT chain advances two links per triple sweep
        if remap and tile >= 6:
            # remapping scheduler: consecutive tiles share the two qubits that sit lowest in the layout between them,
            # layouts chain from pass to pass and the last pass restores the standard order
            std = [n - 1 - q for q in range(n)]
            passes = plan["passes"]
            assert list(passes[-1]["pos_out"]) == std
            for a, b in zip(passes[:-1], passes[1:]):
                assert list(a["pos_out"]) == list(b["pos_in"])
                low = [q for q in range(n) if a["pos_out"][q] < 2]
                assert len(low) == 2 and all(q in a["tile"] and q in b["tile"] for q in low)
            for p in passes:
                assert len(p["tile"]) == tile and len(set(p["tile"])) == tile and sorted(p["pos_out"]) == list(range(n))
        else:
            for p in plan["passes"]:
                assert p["pos_in"] is None
                assert len(p["tile"]) == tile and n - 1 in p["tile"] and n - 2 in p["tile"]
        vals = rng.uniform(-1, 1, size=prog.n_params)
        r = PI.run_devops(plan, vals)
        rho = O.simulate(n, resolve(ops, prog.param_names, vals))
        assert abs(PI.energy(plan, r) - O.energy_sparse(rho, O.pauli_terms_to_matrix(n, *terms))) < TOL


def _clifford_mix_circuit(rng, n, depth):
    """Random mix that exercises every branch of the triple-sweep builder: Clifford gates with Pauli noise (monomial blocks:
    load / store permutations), rotations (chains), dense one- and two-qubit maps, non-unital noise, diagonal two-qubit noise."""
    cz = np.diag([1, 1, 1, -1]).astype(complex)
    swap = np.eye(4)[[0, 2, 1, 3]].astype(complex)
    ops = []
    for _ in range(depth):
        kind = int(rng.integers(0, 10))
        a, b = (int(v) for v in rng.choice(n, 2, replace=False))
        if kind == 0:
            ops += [("u", (a,), O.H), ("k", (a,), O.depolarize(1e-2))]
        elif kind == 1:
            ops += [("u", (a,), O.rx(np.pi / 2 * int(rng.integers(1, 4)))), ("k", (a,), O.bit_flip(2e-2))]
        elif kind in (2, 3):
            ops += [("u", (a, b), O.CNOT), ("k", (a, b), O.two_qubit_depolarize(1e-2))]
        elif kind == 4:
            ops += [("u", (a, b), cz if rng.integers(2) else swap), ("k", (a, b), O.two_qubit_pauli_channel(1e-2, 2e-2, 3e-2))]
        elif kind in (5, 6):
            ops += [("r", (a,), (int(rng.integers(3)), f"t{int(rng.integers(4))}", float(rng.normal()), float(rng.normal()))),
                    ("k", (a,), O.depolarize(1e-2))]
        elif kind == 7:
            ops.append(("k", (a, b), O.two_qubit_depolarize(3e-2)))            # diagonal two-qubit block on its own
        elif kind == 8:
            ops.append(("u", (a, b), random_unitary(rng, 4)))
        else:
            ops.append(("k", (a,), O.amplitude_damp(0.05)))
    return ops


@pytest.mark.parametrize("seed,tile", [(1, 6), (2, 7), (3, 6), (4, 7)])
def test_triple_sweep_lowering_random_clifford_mix(seed, tile):
    n = 8
    rng = np.random.default_rng(800 + seed)
    ops = _clifford_mix_circuit(rng, n, 70)
    terms = random_hamiltonian(rng, n, 12)
    prog = build_program(ops, n, terms, tile_qubits=tile)
    plan = PI.parse(prog.dump())
    assert prog.info["path"] == 2 and plan["sweeps3"] and not plan["sweeps"]
    assert any(sw["lperm"] or sw["sperm"] for sw in plan["sweeps3"])
    vals = rng.uniform(-1, 1, size=prog.n_params)
    rho = O.simulate(n, resolve(ops, prog.param_names, vals))
    want = O.energy_sparse(rho, O.pauli_terms_to_matrix(n, *terms))
    assert abs(PI.energy(plan, PI.run_devops(plan, vals)) - want) < TOL


@pytest.mark.parametrize("n,tile", [(11, 6), (13, 7), (14, 6)])
def test_remap_scheduler_invariants_on_wide_registers(n, tile):
    """Planning only (no state is built): random long-range Clifford + rotation circuits on 11..14 qubits keep the
    scheduler's invariants — every block scheduled once, layouts chain, shared low pair resident on both sides."""
    rng = np.random.default_rng(1400 + n)
    ops = []
    for _ in range(150):
        a, b = (int(v) for v in rng.choice(n, 2, replace=False))
        kind = int(rng.integers(0, 4))
        if kind == 0:
            ops += [("u", (a, b), O.CNOT), ("k", (a, b), O.two_qubit_depolarize(1e-2))]
        elif kind == 1:
            ops += [("r", (a,), (int(rng.integers(3)), f"t{int(rng.integers(6))}", 1.0, 0.0)), ("k", (a,), O.depolarize(1e-2))]
        elif kind == 2:
            ops += [("u", (a,), O.H), ("k", (a,), O.bit_flip(1e-2))]
        else:
            ops.append(("k", (a,), O.amplitude_damp(0.02)))
    terms = random_hamiltonian(rng, n, 7)
    prog = build_program(ops, n, terms, tile_qubits=tile)
    plan = PI.parse(prog.dump())
    assert prog.info["path"] == 2 and prog.info["workspace_bytes_per_state"] == 2 * prog.info["state_bytes"]
    passes = plan["passes"]
    assert list(passes[-1]["pos_out"]) == [n - 1 - q for q in range(n)]
    for a, b in zip(passes[:-1], passes[1:]):
        assert list(a["pos_out"]) == list(b["pos_in"])
        low = [q for q in range(n) if a["pos_out"][q] < 2]
        assert all(q in a["tile"] and q in b["tile"] for q in low)
    assert all(len(p["tile"]) == tile for p in passes)
    assert sum(ph["nsweeps"] for ph in plan["phases"]) == len(plan["sweeps3"])
    # every micro-op of a triple sweep is one of the four triple kinds; chains are referenced exactly once
    assert all(m["kind"] in (16, 17, 18, 19) for m in plan["mus"])
    n_chain_refs = sum(1 for m in plan["mus"] if m["kind"] == 16 and m["a2"])
    assert n_chain_refs == len(plan["chains"])


@pytest.mark.parametrize("seed,tile", [(1, 6), (2, 7), (3, 6), (4, 7), (5, 6)])
def test_wide_monomial_sweeps_random_clifford_mix(seed, tile):
    """Plan option "wide": leading / trailing Clifford x Pauli-noise blocks on ANY tile qubits ride on the exchange addressing
    (thread-base masks) with thread-variant scale tables.  The interpreted program equals the oracle, fewer sweeps are needed
    than with triple-local monomials, and some sweeps really are wide."""
    n = 8
    rng = np.random.default_rng(800 + seed)
    ops = _clifford_mix_circuit(rng, n, 70)
    terms = random_hamiltonian(rng, n, 12)
    local = PI.parse(build_program(ops, n, terms, tile_qubits=tile, wide=0).dump())
    prog = build_program(ops, n, terms, tile_qubits=tile, wide=1)
    plan = PI.parse(prog.dump())
    assert prog.info["path"] == 2 and plan["sweeps3"] and not plan["sweeps"]
    assert any(sw["wide"] for sw in plan["sweeps3"]) and len(plan["sweeps3"]) < len(local["sweeps3"])
    for sw in plan["sweeps3"]:
        assert sw["wide"] in (0, 5, 6, 7) and (not sw["wide"] or (sw["lnvar"] <= 4 and sw["snvar"] <= 4))
    vals = rng.uniform(-1, 1, size=prog.n_params)
    rho = O.simulate(n, resolve(ops, prog.param_names, vals))
    want = O.energy_sparse(rho, O.pauli_terms_to_matrix(n, *terms))
    assert abs(PI.energy(plan, PI.run_devops(plan, vals)) - want) < TOL


def test_wide_monomial_sweeps_hea_layers():
    """The HEA shape of BASELINE configs[3] at 9 qubits / 2 layers: wide sweeps cut the shared-memory round trips."""
    n = 9
    ops = []
    for layer in range(2):
        for q in range(n):
            ops += [("r", (q,), (1, f"y{layer}_{q}", 1.0, 0.0)), ("k", (q,), O.depolarize(1e-3)),
                    ("r", (q,), (2, f"z{layer}_{q}", 1.0, 0.0)), ("k", (q,), O.depolarize(1e-3))]
        for q in range(n - 1):
            ops += [("u", (q, q + 1), O.CNOT), ("k", (q, q + 1), O.two_qubit_depolarize(1e-3))]
    rng = np.random.default_rng(2)
    terms = random_hamiltonian(rng, n, 8)
    counts = {}
    for wide in (0, 1):
        prog = build_program(ops, n, terms, tile_qubits=6, wide=wide)
        plan = PI.parse(prog.dump())
        counts[wide] = len(plan["sweeps3"])
        vals = rng.uniform(-1, 1, size=prog.n_params)
        rho = O.simulate(n, resolve(ops, prog.param_names, vals))
        assert abs(PI.energy(plan, PI.run_devops(plan, vals)) - O.energy_sparse(rho, O.pauli_terms_to_matrix(n, *terms))) < TOL
    assert counts[1] < counts[0]
