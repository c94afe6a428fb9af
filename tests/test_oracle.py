"""The oracle against everything the reference pins — CPU only."""
import json
import os

import numpy as np
import pytest
from scipy.optimize import minimize_scalar

from oracle import c_oracle as CO
from oracle import dm_oracle as O
from workloads import ALPHA_STAR, GOLDEN, h2_golden, h2_ops, h2_terms
from vqe_errormitigation_b200 import engine


def h2_energy(alpha, terms_matrix, n1=(), n2=()):
    return O.energy_sparse(O.simulate(4, h2_ops(list(n1), list(n2), alpha=alpha)), terms_matrix)


@pytest.mark.parametrize("r", ["0.3", "0.7414", "1.5", "3.0"])
def test_hf_and_fci_known_answers(r):
    """Noiseless E(0) = stored hf_energy; min_alpha E = lambda_min(H) = stored fci_energy (reference hdf5 values)."""
    rec = h2_golden()["molecules"][r]
    ham = O.pauli_terms_to_matrix(4, *h2_terms(r))
    assert abs(h2_energy(0.0, ham) - rec["hf_energy"]) < 1e-12
    assert abs(np.linalg.eigvalsh(ham)[0] - rec["fci_energy"]) < 1e-12
    res = minimize_scalar(lambda a: h2_energy(a, ham), bracket=(-0.3, 0.3), tol=1e-12)
    assert abs(res.fun - rec["fci_energy"]) < 1e-10


def test_all_bond_lengths_hamiltonian_known_answers():
    for r, rec in h2_golden()["molecules"].items():
        ham = O.pauli_terms_to_matrix(4, *h2_terms(r))
        assert abs(ham[12, 12].real - rec["hf_energy"]) < 1e-12, r     # <1100|H|1100>
        assert abs(np.linalg.eigvalsh(ham)[0] - rec["fci_energy"]) < 1e-12, r


def test_alpha_star():
    ham = O.pauli_terms_to_matrix(4, *h2_terms())
    assert abs(h2_energy(ALPHA_STAR, ham) - h2_golden()["molecules"]["0.7414"]["fci_energy"]) < 1e-12


def test_golden_noisy_energies_regression():
    with open(os.path.join(GOLDEN, "h2_noisy_energies.json")) as fh:
        gold = json.load(fh)
    ham = O.pauli_terms_to_matrix(4, *h2_terms())
    want = gold["energies"]["depolarize1e-3+2qdepolarize1e-3"]
    for a, e in zip(gold["alphas"], want):
        assert abs(h2_energy(a, ham, [O.depolarize(1e-3)], [O.two_qubit_depolarize(1e-3)]) - e) < 1e-13
    # values the survey derived independently, to its 12 printed digits
    assert abs(want[0] - (-1.038084151862)) < 1e-11 and abs(want[1] - (-1.055947703192)) < 1e-11


def test_complex64_mode_reproduces_reference_precision():
    ham = O.pauli_terms_to_matrix(4, *h2_terms())
    ops = h2_ops([O.depolarize(1e-3)], [O.two_qubit_depolarize(1e-3)], alpha=ALPHA_STAR)
    e64 = O.energy_sparse(O.simulate(4, ops), ham)
    e32 = O.energy_sparse(O.simulate(4, ops, dtype=np.complex64).astype(np.complex128), ham)
    assert 1e-9 < abs(e64 - e32) < 1e-4


def test_quasi_probability_known_answers():
    """Analytic values of SURVEY.md App. C.2."""
    q = O.quasi_probabilities(O.H, [O.depolarize(1e-3)])
    assert abs(q[0] - 1.0010013351134845) < 1e-13
    assert np.allclose(q[1:4], -3.3377837116155273e-4, atol=1e-15)
    assert np.allclose(q[4:], 0, atol=1e-14)
    q = O.quasi_probabilities(O.rx(0.3), [O.bit_flip(1e-3)])
    assert abs(q[0] - 1.001002004008016) < 1e-13 and abs(q[1] + 0.001002004008016) < 1e-13
    q2 = O.quasi_probabilities(O.CNOT, [O.two_qubit_depolarize(1e-3)])
    assert abs(np.sum(np.abs(q2)) - 1.0020021356113191) < 1e-12
    assert abs(q2[0] - 1.001001068) < 1e-8
    q1 = O.quasi_probabilities(O.H, [O.asymmetric_depolarize(1e-4, 1e-4, 6e-4)])
    assert abs(np.sum(np.abs(q1)) - 1.00160204277986) < 1e-12
    q2 = O.quasi_probabilities(O.CNOT, [O.two_qubit_pauli_channel(1e-4, 1e-4, 6e-4)])
    assert abs(np.sum(np.abs(q2)) - 1.003206652100788) < 1e-12


def test_basis_operations_table():
    b = O.basis_operations()
    assert np.allclose(b[1], 1j * O.X) and np.allclose(b[2], 1j * O.Y) and np.allclose(b[3], 1j * O.Z)
    assert np.allclose(b[12], [[1, 0], [0, 0]]) and np.allclose(b[10], 0.5 * np.ones((2, 2)))
    assert np.allclose(b[15], [[0, 1j], [0, 0]])
    for k in range(10):
        assert np.allclose(b[k] @ b[k].conj().T, np.eye(2))
    # the QP solve is invertible: reconstruct N^-1 from the coefficients
    q = O.quasi_probabilities(O.H, [O.depolarize(0.01)])
    rec = sum(c * O.pauli_transfer_matrix(m) for c, m in zip(q, b))
    clean, noisy = O.pauli_transfer_matrix(O.H), O.pauli_transfer_matrix(O.H, O.kraus_on_matrix(O.depolarize(0.01)))
    assert np.allclose(rec, clean @ np.linalg.inv(noisy), atol=1e-12)


def test_channel_closed_forms_and_invariants():
    """App. E identities: trace preservation, Hermiticity, TwoQubitPauliChannel = product of 1q channels,
    TwoQubitDepolarizingChannel closed form."""
    rng = np.random.default_rng(3)
    m = rng.normal(size=(16, 16)) + 1j * rng.normal(size=(16, 16))
    rho = m @ m.conj().T
    rho /= np.trace(rho)
    t = rho.reshape((2,) * 8)
    for kraus, qs in ((O.depolarize(0.1), (2,)), (O.two_qubit_depolarize(0.2), (0, 3)), (O.amplitude_damp(0.3), (1,)),
                      (O.two_qubit_pauli_channel(0.01, 0.02, 0.03), (1, 2))):
        out = O.apply_op(t, 4, "k", qs, kraus).reshape(16, 16)
        assert abs(np.trace(out) - 1) < 1e-13
        assert np.allclose(out, out.conj().T, atol=1e-14)
    a = O.apply_op(t, 4, "k", (1, 2), O.two_qubit_pauli_channel(0.01, 0.02, 0.03))
    b = O.apply_op(O.apply_op(t, 4, "k", (1,), O.asymmetric_depolarize(0.01, 0.02, 0.03)), 4, "k", (2,), O.asymmetric_depolarize(0.01, 0.02, 0.03))
    assert np.allclose(a, b, atol=1e-15)
    p = 0.2
    out = O.apply_op(t, 4, "k", (0, 3), O.two_qubit_depolarize(p)).reshape(16, 16)
    # (1 - 16p/15) rho + (4p/15) I/4... : partial trace over qubits (0,3) tensored with identity
    t8 = rho.reshape((2,) * 8)
    red = np.einsum("abcdaecd->be" if False else "abcdefgh->abcdefgh", t8)  # keep shape; explicit loop below
    closed = (1 - 16 * p / 15) * rho.copy()
    for r in range(16):
        for c in range(16):
            mask = 0b1001
            if (r & mask) == (c & mask):
                closed[r, c] += (4 * p / 15) * sum(rho[(r & ~mask) | v, (c & ~mask) | v] for v in (0, 1, 8, 9))
    assert np.allclose(out, closed, atol=1e-14)


def _builder(ops, n):
    b = engine.ProgramBuilder(n)
    for kind, qs, pl in ops:
        if kind == "u":
            b.add_unitary(qs, pl)
        elif kind == "k":
            b.add_kraus(qs, pl)
        elif kind == "r":
            b.add_rotation(qs[0], *pl)
        elif kind == "v":
            b.add_variant(qs, pl[0], slot=pl[1])
        elif kind == "ck":
            b.add_kraus(qs, pl[0], if_slot=pl[1])
    return b


def test_c_oracle_matches_numpy_oracle():
    terms = h2_terms()
    ham = O.pauli_terms_to_matrix(4, *terms)
    n1, n2 = [O.depolarize(1e-3)], [O.two_qubit_depolarize(1e-3)]
    b = _builder(h2_ops(n1, n2), 4)
    alphas = np.array([[0.0], [ALPHA_STAR], [-0.37], [2.5]])
    got = CO.energy_batch(4, b.ops, b.pool(), terms, params=alphas, n_params=1, threads=2)
    for a, g in zip(alphas[:, 0], got):
        assert abs(g - h2_energy(a, ham, n1, n2)) < 1e-13
    rho = CO.density(4, b.ops, b.pool(), params=alphas[1])
    assert np.max(np.abs(rho - O.simulate(4, h2_ops(n1, n2, alpha=ALPHA_STAR)))) < 1e-14


def test_c_oracle_variants_match_numpy_oracle():
    with open(os.path.join(GOLDEN, "h2_noisy_energies.json")) as fh:
        gold = json.load(fh)
    terms = h2_terms()
    n1, n2 = [O.bit_flip(1e-3), O.depolarize(1e-3)], [O.two_qubit_depolarize(1e-3)]
    basis = O.basis_operations()
    ops = []
    for s, (kind, qs, pl) in enumerate(O.h2_gate_list(ALPHA_STAR)):
        noise = n1 if len(qs) == 1 else n2
        ops.append((kind, qs, pl))
        ops += [("k", qs, kr) for kr in noise]
        ops.append(("v", qs, (basis, s)))
        ops += [("ck", qs, (kr, s)) for kr in noise]
    b = _builder(ops, 4)
    rows = np.array([c["identifier"] for c in gold["variant_cases"]], np.uint8)
    got = CO.energy_batch(4, b.ops, b.pool(), terms, variants=rows, n_slots=122, threads=2)
    want = np.array([c["energy"] for c in gold["variant_cases"]])
    assert np.max(np.abs(got - want)) < 1e-13


def test_swap_test_noisy_value_matches_reference_stored_run():
    """The only reference-GENERATED noisy value without an optimiser in the loop: the 3+3-qubit SWAP test under
    SingleQubit/TwoQubitPauliChannel(1e-4, 1e-4, 6e-4), 100 repetitions of 10 000 shots
    (src/visual_testing.py:492-524 -> src/classic_minimisation/SWAP_Similarity_P0.0001_R100_C10000).  Its expectation
    is <Z0> of the noisy circuit; the oracle's exact value must sit inside the stored run's standard error."""
    import json
    import os
    from vqe_errormitigation_b200 import cirq_compat as cirq
    from vqe_errormitigation_b200.qp_qem_lib import swap_circuit
    q = cirq.LineQubit.range(7)
    n, ops_all = O.from_circuit(swap_circuit(3, q, True), q)
    gates = [g for g in ops_all if g[0] == "u"]          # get_updated_circuit puts no noise after the measurement (:540-543)
    assert sum(len(g[1]) == 2 for g in gates) == 38 and sum(len(g[1]) == 1 for g in gates) == 39   # 3 x 24 + 5
    p = 1e-4
    ops = O.with_noise(gates, [O.asymmetric_depolarize(p, p, 6 * p)], [O.two_qubit_pauli_channel(p, p, 6 * p)])
    probs = O.probabilities(O.simulate(n, ops))
    z0 = float(probs[:64].sum() - probs[64:].sum())
    assert abs(z0 - 0.45602390888719) < 1e-12
    with open(os.path.join(os.path.dirname(__file__), "golden", "classic_minimisation.json")) as fh:
        rec = json.load(fh)["files"]["SWAP_Similarity_P0.0001_R100_C10000"]["data_noisy"]
    standard_error = rec["std"] / np.sqrt(rec["n"])
    assert abs(z0 - rec["mean"]) < standard_error          # 0.456024 vs 0.456066 +- 0.00085
    ideal = O.probabilities(O.simulate(n, gates))
    assert abs(float(ideal[:64].sum() - ideal[64:].sum()) - 0.5) < 1e-13


def _embed(n, qubits, m):
    """Full 2^n x 2^n matrix of a k-qubit operator on ``qubits`` (qubit 0 = most significant bit), element by element: an
    index-arithmetic construction that shares nothing with the oracle's tensordot / moveaxis path."""
    k, dim = len(qubits), 1 << n
    shifts = [n - 1 - q for q in qubits]
    clear = dim - 1
    for s in shifts:
        clear &= ~(1 << s)
    full = np.zeros((dim, dim), dtype=complex)
    for col in range(dim):
        sub_c = sum(((col >> s) & 1) << (k - 1 - i) for i, s in enumerate(shifts))
        for sub_r in range(1 << k):
            if m[sub_r, sub_c] == 0:
                continue
            row = col & clear
            for i, s in enumerate(shifts):
                row |= ((sub_r >> (k - 1 - i)) & 1) << s
            full[row, col] = m[sub_r, sub_c]
    return full


def test_noisy_energies_by_an_independent_liouville_restatement():
    """Third arithmetic path for the noisy values no reference fixture pins: the whole circuit as ONE 256 x 256
    superoperator S = prod_ops sum_k K (x) conj(K) acting on the row-major vec(rho) - no per-op density-matrix update at all -
    against the per-op oracle, the stored golden energies and SURVEY App. A.4's values."""
    ham = O.pauli_terms_to_matrix(4, *h2_terms())
    cases = [((O.depolarize(1e-3),), (O.two_qubit_depolarize(1e-3),), ALPHA_STAR, -1.0559477031919418),
             ((O.depolarize(1e-3),), (O.two_qubit_depolarize(1e-3),), 0.0, -1.0380841518615922),
             ((O.bit_flip(1e-3), O.depolarize(1e-3)), (O.two_qubit_depolarize(1e-3),), ALPHA_STAR, None),
             ((O.amplitude_damp(0.01),), (O.two_qubit_pauli_channel(1e-4, 1e-4, 6e-4),), 0.3, None)]
    for n1, n2, alpha, known in cases:
        ops = h2_ops(list(n1), list(n2), alpha=alpha)
        total = np.eye(256, dtype=complex)
        for kind, qubits, payload in ops:
            kraus = [payload] if kind == "u" else payload
            step = sum(np.kron(f, f.conj()) for f in (_embed(4, qubits, np.asarray(kr, complex)) for kr in kraus))
            total = step @ total
        vec0 = np.zeros(256, dtype=complex)
        vec0[0] = 1.0
        rho = (total @ vec0).reshape(16, 16)
        ref = O.simulate(4, ops)
        assert np.max(np.abs(rho - ref)) < 1e-13
        energy = float(np.real(np.trace(rho @ ham)))
        assert abs(energy - O.energy_sparse(ref, ham)) < 1e-12
        if known is not None:
            assert abs(energy - known) < 1e-12
    # a quasi-probability variant row with every kind of basis operation (Pauli, R-type, projector; 1- and 2-qubit)
    rng = np.random.default_rng(11)
    gates = O.h2_gate_list(ALPHA_STAR)
    identifier = [int(rng.integers(0, 16 if len(q) == 1 else 256)) if rng.random() < 0.15 else 0 for _, q, _ in gates]
    assert sum(1 for i in identifier if i) >= 10
    ops = O.variant_ops(gates, [O.depolarize(1e-3)], [O.two_qubit_depolarize(1e-3)], identifier)
    vec = np.zeros(256, dtype=complex)
    vec[0] = 1.0
    for kind, qubits, payload in ops:
        kraus = [payload] if kind == "u" else payload
        vec = sum(np.kron(f, f.conj()) for f in (_embed(4, qubits, np.asarray(kr, complex)) for kr in kraus)) @ vec
    ref = O.simulate(4, ops)
    assert np.max(np.abs(vec.reshape(16, 16) - ref)) < 1e-13 * max(1.0, np.max(np.abs(ref)))
