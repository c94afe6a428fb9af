"""Device-side QP variant sampler (dmx_qp_sample_variants) — the device twin of QuasiCircuitIdentifier._get_identifiers
(reference src/error_mitigation.py:309-335).  CPU: the numpy restatement of its RNG against the published Philox4x32-10
known answers.  GPU: bit-exact equality with that restatement, the sampled distribution, the sign rule, rank slicing and
the device-resident mitigated estimator."""
import numpy as np
import pytest

import philox_ref as P


def test_philox_reference_known_answers():
    """Random123 kat_vectors for philox4x32-10 (counter, key -> output)."""
    kat = [((0, 0, 0, 0), 0, (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, 0xffffffffffffffff, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0x299f31d0 << 32) | 0xa4093822, (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = P.philox4x32_10(*[[c] for c in ctr], key)
        assert tuple(int(g[0]) for g in got) == want


def test_reference_sampler_distribution_and_sign_rule():
    qps = [np.array([1.2, -0.1, 0.0, -0.1]), np.array([0.5, 0.5]), np.linspace(-1, 1, 16)]
    rows, signs = P.sample_variants(qps, 200000, seed=7)
    for s, qp in enumerate(qps):
        p = np.abs(qp) / np.abs(qp).sum()
        freq = np.bincount(rows[:, s], minlength=len(qp)) / rows.shape[0]
        assert np.max(np.abs(freq - p)) < 5e-3
        assert freq[p == 0].sum() == 0                      # zero-probability entries are never drawn
    want = np.prod([np.sign(qp[rows[:, s]]) for s, qp in enumerate(qps)], axis=0)
    assert np.array_equal(signs, want)
    # slices of one stream: first_sample offsets reproduce the same samples
    rows2, signs2 = P.sample_variants(qps, 1000, seed=7, first_sample=5000)
    assert np.array_equal(rows2, rows[5000:6000]) and np.array_equal(signs2, signs[5000:6000])


@pytest.mark.gpu
def test_device_sampler_is_bit_exact_with_reference():
    import torch
    from vqe_errormitigation_b200 import engine
    rng = np.random.default_rng(11)
    qps = [rng.normal(size=k) for k in (16, 256, 16, 2, 1, 256, 16)] + [np.array([1.3, -0.1, -0.1, -0.1] + [0.0] * 12)]
    odd = [np.array([0.7, -0.3]), np.array([1.0]), np.array([-2.0, 1.0, 0.5])]      # 3 slots x stride 3: sub-tables need padding to 16 B
    # 150 slots of near-identity tables (the QEM regime: the whole-group fast test decides almost every lane), some with a
    # NEGATIVE or ZERO-weight identity entry (sign summary; threshold 0 = the identity is never drawn), more than 128 slots
    wide = []
    for k in range(150):
        t = np.array([1.0 + 3e-3, -1e-3, -1e-3, -1e-3])
        if k % 7 == 3:
            t = -t
        if k % 31 == 5:
            t = np.array([0.0, 0.6, -0.4, 0.2])
        wide.append(t)
    for tables in (qps, odd, odd[:1], qps[:5], wide, wide[:122], wide[:129]):
        for batch, seed, first in ((1, 0, 0), (4097, 123456789012345, 0), (1000, 5, (1 << 33) + 17)):
            dv, ds = engine.qp_sample_variants(tables, batch, seed, first_sample=first)
            rv, rs = P.sample_variants(tables, batch, seed, first_sample=first)
            assert np.array_equal(dv.cpu().numpy(), rv)
            assert np.array_equal(ds.cpu().numpy(), rs)


@pytest.mark.gpu
def test_device_resident_mitigation_estimator():
    """get_mu_effective_sampled: sampler + batched variant evaluation + reduction on the device equal the same estimator
    assembled on the host from the reference-sampler's rows, and the mitigated value moves towards the noiseless one."""
    import torch
    from vqe_errormitigation_b200 import cirq_compat as cirq, engine
    from vqe_errormitigation_b200.data_containers.helper_interfaces.i_noise_wrapper import INoiseModel, INoiseWrapper
    from vqe_errormitigation_b200.data_containers.model_hydrogen import HydrogenAnsatz
    from vqe_errormitigation_b200.error_mitigation import IErrorMitigationManager, TwoQubitDepolarizingChannel
    from vqe_errormitigation_b200.processors.processor_quantum import QPU
    w = HydrogenAnsatz()
    w.operator_parameters["alpha0"] = 0.014133516193216759
    noise = INoiseModel([cirq.depolarize(2e-3)], [TwoQubitDepolarizingChannel(2e-3)], "d")
    n_w = INoiseWrapper(w, noise)
    clean = QPU.get_resolved_circuit(n_w.get_clean_circuit(), w.operator_parameters)
    mgr = IErrorMitigationManager(clean, noise, QPU.get_hamiltonian_objective_operator(w))
    N, seed = 40000, 20260102
    mean, stderr, n = mgr.get_mu_effective_sampled(N, seed)
    assert n == N
    qps = mgr.quasi_probabilities()
    rows, signs = P.sample_variants(qps, N, seed)
    scale = float(np.prod([np.sum(np.abs(q)) for q in qps]))
    values = mgr.variant_program().energies(None, torch.tensor(rows)).cpu().numpy()
    want = np.mean(signs * scale * values)
    assert abs(mean - want) < 1e-10
    assert abs(stderr - np.std(signs * scale * values) / np.sqrt(N)) < 1e-9
    ideal = QPU.get_simulated_noisy_expectation_value(w, clean, cirq.ParamResolver({}))
    noisy = QPU.get_simulated_noisy_expectation_value(n_w, QPU.get_resolved_circuit(n_w.get_noisy_circuit(), w.operator_parameters), cirq.ParamResolver({}))
    assert abs(mean - ideal) < 6 * stderr + 1e-4 and abs(mean - ideal) < abs(noisy - ideal)


@pytest.mark.gpu
def test_fused_draw_and_evaluate_is_bit_identical_to_the_two_step_path():
    """dmx_qp_sampled_energies (dmx_frame16q_kernel: one thread draws AND evaluates one sample) against QpSampler.draw +
    CircuitProgram.energies on the same (seed, first_sample): signs and energies bit for bit - with the circuit's own
    quasi-probabilities, with tables that put weight on R-type / projector operations (those samples are re-drawn into a
    compact table for the general tile path), with tables dense enough to overflow a thread's entry list, on ragged batches
    and stream offsets; and against the numpy Philox restatement + the C oracle for a few samples."""
    import torch
    from oracle import dm_oracle as O
    from test_gpu_parity import oracle_energy, qp_noisy_variant_program
    from workloads import h2_terms
    from vqe_errormitigation_b200 import engine
    terms = h2_terms()
    ham = O.pauli_terms_to_matrix(4, *terms)
    n1, n2 = [O.bit_flip(1e-3), O.depolarize(1e-3)], [O.two_qubit_depolarize(1e-3)]
    prog, gates = qp_noisy_variant_program(n1, n2, terms)
    assert " const=1" in prog.dump()
    arity = [len(g[1]) for g in gates]
    rng = np.random.default_rng(99)

    def tables(kind):
        out = []
        for a in arity:
            k = 16 if a == 1 else 256
            t = np.zeros(k)
            t[0] = 1.0 + 4e-3
            pauli = [1, 2, 3] if a == 1 else [16 * i + j for i in range(4) for j in range(4) if i or j]
            for j in pauli:
                t[j] = -4e-3 / len(pauli) * rng.uniform(0.5, 1.5)
            if kind == "rtype":                       # weight on rotations / projectors: general path
                t[7 if a == 1 else 7 * 16 + 2] = 3e-3
            if kind == "dense":                       # ~8 non-identity draws per row: the entry list overflows for some rows
                t[0] = 1.0
                for j in pauli:
                    t[j] = 0.07 / len(pauli)
            out.append(t)
        return out

    launches = {}
    for kind in ("pauli", "rtype", "dense"):
        qps = tables(kind)
        sampler = engine.QpSampler(qps)
        n0 = engine.launch_count()
        sampler.sampled_energies(prog, 4099, 1)
        launches[kind] = engine.launch_count() - n0
        for batch, seed, first in ((4099, 7, 0), (2048, 123456789012345, (1 << 33) + 5), (130, 3, 17)):
            values, signs = sampler.sampled_energies(prog, batch, seed, first_sample=first)
            rows, signs2 = sampler.draw(batch, seed, first_sample=first)
            want = prog.energies(None, rows)
            assert torch.equal(signs, signs2), (kind, batch)
            if batch >= 2048 and kind != "dense":     # the two-step path runs the same thread-per-row arithmetic
                assert torch.equal(values, want), (kind, batch, float((values - want).abs().max()))
            elif batch >= 2048:                       # rows with more than 6 entries: local-memory list here, tile kernel there
                assert float((values - want).abs().max()) < 1e-13
                assert float((values == want).double().mean()) > 0.2          # the rows that fit the list are bit-identical
            else:                                     # ... and the warp-per-row kernel below 2048 rows: equal to rounding
                assert float((values - want).abs().max()) < 1e-13
        ref_rows, ref_signs = P.sample_variants(qps, 64, 7)
        values, signs = sampler.sampled_energies(prog, 4099, 7)
        assert np.array_equal(signs[:64].cpu().numpy(), ref_signs)
        for b in range(0, 64, 7):
            assert abs(float(values[b]) - oracle_energy(O.variant_ops(gates, n1, n2, ref_rows[b]), 4, ham)) < 1e-10, (kind, b)
    # one launch when no drawable operation can need the general path, three (kernel, re-draw of the listed samples, tile kernel) otherwise
    assert launches == {"pauli": 1, "rtype": 3, "dense": 1}


def test_reference_shot_sampler_statistics():
    """numpy restatement of the device shot sampler (philox_ref.shot_expectations): parity expectations converge to the exact
    values of the distribution, every mask is read from the same shots, rows use independent streams."""
    rng = np.random.default_rng(3)
    probs = rng.dirichlet(np.ones(16), size=3)
    masks = [0b1000, 0b0100, 0b1100, 0b1111]
    shots = 200000
    got = P.shot_expectations(probs, shots, masks, seed=5)
    for r in range(3):
        for j, m in enumerate(masks):
            sign = np.array([1 - 2 * (bin(i & m).count("1") & 1) for i in range(16)])
            assert abs(got[r, j] - float(np.dot(sign, probs[r]))) < 5 / np.sqrt(shots)
    # a deterministic distribution gives exact parities; first_row shifts the stream; per-row shot counts
    delta = np.zeros((1, 16)); delta[0, 0b1010] = 1.0
    assert P.shot_expectations(delta, 17, masks, seed=1).tolist() == [[-1.0, 1.0, -1.0, 1.0]]
    a = P.shot_expectations(probs, 1000, masks, seed=9, first_row=1)
    b = P.shot_expectations(probs[:1], 1000, masks, seed=9, first_row=1)
    c = P.shot_expectations(np.concatenate([probs[:1], probs]), 1000, masks, seed=9)
    assert np.array_equal(a[:1], b) and np.array_equal(c[1], b[0])       # row 1 of c: probs[0] on stream row 1, like b
    d = P.shot_expectations(probs, np.array([10, 1000, 7]), masks, seed=9)
    assert np.array_equal(d[1], P.shot_expectations(probs[1:2], 1000, masks, seed=9, first_row=1)[0])


@pytest.mark.gpu
def test_device_shot_sampler_is_bit_exact_with_reference():
    import torch
    from vqe_errormitigation_b200 import engine
    rng = np.random.default_rng(21)
    for dim, masks in ((16, [0b1000, 0b0001, 0b0110, 0b1111]), (128, [1, 64, 127, 0b1010101]), (2, [1])):
        probs = rng.dirichlet(np.ones(dim) * 0.5, size=37)
        probs[3] = 0.0
        probs[3, dim - 1] = 1.0
        for shots in (1, 10, 1001):
            want = P.shot_expectations(probs, shots, masks, seed=1234567, first_row=5)
            got = engine.shot_expectations(torch.tensor(probs, device="cuda"), shots, masks, 1234567, first_row=5).cpu().numpy()
            assert np.array_equal(got, want), (dim, shots)
        per_row = rng.integers(1, 500, size=37)
        want = P.shot_expectations(probs, per_row, masks, seed=99)
        got = engine.shot_expectations(torch.tensor(probs, device="cuda"), per_row, masks, 99).cpu().numpy()
        assert np.array_equal(got, want)


@pytest.mark.gpu
def test_sampled_energy_program_batched_matches_exact_limit():
    """SampledEnergyProgram (the five measurement circuits of measure_observable as one plan, shots on the device): the
    batched sampled energies of several parameter rows converge to the exact sum_t w_t <P_t> of their own circuits, rows
    are independent, and a fixed seed reproduces the numbers."""
    import torch
    from oracle import dm_oracle as O
    from vqe_errormitigation_b200 import cirq_compat as cirq
    from vqe_errormitigation_b200.data_containers.helper_interfaces.i_noise_wrapper import INoiseModel, INoiseWrapper
    from vqe_errormitigation_b200.data_containers.helper_interfaces.i_wave_function import IGeneralizedUCCSD
    from vqe_errormitigation_b200.data_containers.model_hydrogen import HydrogenAnsatz, SampledEnergyProgram
    from vqe_errormitigation_b200.error_mitigation import TwoQubitDepolarizingChannel
    from vqe_errormitigation_b200.openfermion_compat import jordan_wigner
    from vqe_errormitigation_b200.processors.processor_quantum import QPU
    w = HydrogenAnsatz()
    model = INoiseModel([cirq.depolarize(2e-3)], [TwoQubitDepolarizingChannel(2e-3)], "d")
    n_w = INoiseWrapper(w, model)
    ham = jordan_wigner(w.molecule.get_molecular_hamiltonian())
    prog = SampledEnergyProgram(n_w.get_noisy_circuit(), w.qubits, ham, model.get_operators, IGeneralizedUCCSD.pauli_basis_map())
    assert prog.variant_rows.shape == (5, 4) and prog.prog.param_names == ["alpha0"]
    alphas = np.array([[-0.2], [0.0141], [0.3]])
    shots = 400000
    got = prog.energies(alphas, shots, seed=11)
    assert np.array_equal(got, prog.energies(alphas, shots, seed=11)) and not np.array_equal(got, prog.energies(alphas, shots, seed=12))
    basis = w.pauli_basis_map()
    for b, alpha in enumerate(alphas[:, 0]):
        w.operator_parameters["alpha0"] = alpha
        base = QPU.get_resolved_circuit(n_w.get_noisy_circuit(), w.operator_parameters)
        exact = 0.0
        for term, coeff in ham.terms.items():
            if len(term) == 0:
                exact += complex(coeff).real
                continue
            c = base.copy()
            if len(term) == 4:
                c.append([(basis[p](w.qubits[q], 1), model.get_operators(w.qubits[q])) for q, p in term])
            else:
                for q in range(4):
                    c.append([cirq.I(w.qubits[q]), model.get_operators(w.qubits[q])])
            rho = O.simulate(4, O.from_circuit(c, w.qubits)[1])
            zmask = sum(1 << (3 - q) for q, _ in term)
            par = np.array([bin(i & zmask).count("1") % 2 for i in range(16)])
            exact += complex(coeff).real * float(np.sum(np.where(par == 0, 1, -1) * O.probabilities(rho)))
        assert abs(got[b] - exact) < 6 * 0.6 / np.sqrt(shots), (b, got[b], exact)
