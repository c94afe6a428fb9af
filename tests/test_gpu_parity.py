"""GPU parity tests: CUDA path (through the C ABI) vs the numpy oracle on the same seeded inputs.

Tolerance: 1e-10 absolute on energies / density-matrix elements / probabilities — the bar BASELINE.json's
north_star states for FP64.  All tests here need a B200 (``-m gpu``).
"""
import numpy as np
import pytest

from oracle import dm_oracle as O
from workloads import ALPHA_STAR, build_program, h2_ops, h2_program, h2_terms

pytestmark = pytest.mark.gpu
TOL = 1e-10


def torch_cuda():
    import torch
    assert torch.cuda.is_available()
    return torch


def oracle_energy(ops, n, ham):
    return O.energy_sparse(O.simulate(n, ops), ham)


NOISES = {
    "depol+2qdepol": ([O.depolarize(1e-3)], [O.two_qubit_depolarize(1e-3)]),
    "depol_only_1q": ([O.depolarize(1e-3)], []),
    "asym_pauli": ([O.asymmetric_depolarize(1e-4, 1e-4, 6e-4)], [O.two_qubit_pauli_channel(1e-4, 1e-4, 6e-4)]),
    "bitflip+phaseflip": ([O.bit_flip(1e-3), O.phase_flip(2e-3)], []),
    "noiseless": ([], []),
    "amp_damp": ([O.amplitude_damp(5e-3)], [O.two_qubit_depolarize(1e-3)]),
}


@pytest.mark.parametrize("noise", sorted(NOISES))
@pytest.mark.parametrize("options", [{}, {"segments": 0}, {"segments": 0, "fuse": 0}])
def test_h2_energy_matches_oracle(noise, options):
    torch = torch_cuda()
    n1, n2 = NOISES[noise]
    terms = h2_terms()
    ham = O.pauli_terms_to_matrix(4, *terms)
    prog = build_program(h2_ops(n1, n2), 4, terms, **options)
    if noise == "amp_damp":
        assert prog.info["path"] == 0  # non-Pauli channel: dense blocks, tile kernel
    elif not options:
        assert prog.info["path"] == 1
    alphas = np.array([0.0, ALPHA_STAR, -0.37, 0.5, 1.234, -2.0, 3.0])
    got = prog.energies(torch.tensor(alphas).reshape(-1, 1)).cpu().numpy()
    for a, g in zip(alphas, got):
        want = oracle_energy(h2_ops(n1, n2, alpha=a), 4, ham)
        assert abs(g - want) < TOL, (noise, options, a, g, want)


def test_h2_known_answers_all_bond_lengths():
    """Noiseless E(alpha=0) = stored hf_energy for all 30 bond lengths; min over alpha reaches fci_energy."""
    torch = torch_cuda()
    from workloads import h2_golden
    alphas = np.linspace(-0.6, 0.6, 4801)
    for r, rec in h2_golden()["molecules"].items():
        terms = h2_terms(r)
        prog = build_program(h2_ops([], []), 4, terms)
        e = prog.energies(torch.tensor(alphas).reshape(-1, 1)).cpu().numpy()
        e0 = prog.energies(torch.zeros(1, 1, dtype=torch.float64)).cpu().numpy()[0]
        assert abs(e0 - rec["hf_energy"]) < 1e-10, r
        # grid minimum is within the quadratic error of the grid step of the FCI energy
        assert 0 <= e.min() - rec["fci_energy"] < 2e-6, (r, e.min(), rec["fci_energy"])


def test_batch_sizes_and_host_api():
    torch = torch_cuda()
    prog, ham = h2_program()
    rng = np.random.default_rng(5)
    for B in (1, 7, 8, 9, 255, 4099):
        a = rng.normal(0, 0.3, size=(B, 1))
        dev = prog.energies(torch.tensor(a)).cpu().numpy()
        host = prog.energies_host(a)
        assert np.max(np.abs(dev - host)) < 1e-13
        i = B // 2
        want = oracle_energy(h2_ops([O.depolarize(1e-3)], [O.two_qubit_depolarize(1e-3)], alpha=a[i, 0]), 4, ham)
        assert abs(dev[i] - want) < TOL
    assert prog.energies(torch.zeros(0, 1, dtype=torch.float64)).numel() == 0
    # host-buffer call, pipelined in chunks (three streams): ragged batch, pageable and page-locked buffers
    from vqe_errormitigation_b200 import engine
    B = 70001
    a = rng.normal(0, 0.3, size=(B, 1))
    dev = prog.energies(torch.tensor(a)).cpu().numpy()
    assert np.max(np.abs(prog.energies_host(a) - dev)) < 1e-13
    ap, out = engine.pinned_empty(a.shape), engine.pinned_empty(B)
    ap[...] = a
    out[...] = np.nan
    res = prog.energies_host(ap, out=out)            # page-locked in and out: zero-copy (the kernel works on the mapped buffers)
    # (mapped buffers run the warp-per-circuit kernel, device-resident batches of this size the thread-per-circuit one: the
    # energy sums differ in their order, so equality holds to rounding; with both on one kernel it is exact - below)
    assert res is out and np.max(np.abs(out - dev)) < 1e-14
    out[...] = np.nan                                 # page-locked out, pageable in: DMA pipeline
    assert np.array_equal(prog.energies_host(a, out=out), dev)
    dma, _ = h2_program(host_zero_copy=0)             # the option is library-wide: restore it before leaving
    try:
        out[...] = np.nan
        assert np.array_equal(dma.energies_host(ap, out=out), dev)
    finally:
        h2_program(host_zero_copy=1)
    for small in (1, 7):                              # ragged CTA tails through the zero-copy path
        sp, so = engine.pinned_empty((small, 1)), engine.pinned_empty(small)
        sp[...] = a[:small]
        assert np.max(np.abs(prog.energies_host(sp, out=so) - dev[:small])) < 1e-14
    both, _ = h2_program(frame16t=2)                  # thread-per-circuit kernel on mapped buffers too: bit-identical rows
    out[...] = np.nan
    assert np.array_equal(both.energies_host(ap, out=out), both.energies(torch.tensor(a)).cpu().numpy())
    assert np.array_equal(out, dev)
    warp, _ = h2_program(frame16t=0)                  # and with the warp-per-circuit kernels on both sides
    out[...] = np.nan
    assert np.array_equal(warp.energies_host(ap, out=out), warp.energies(torch.tensor(a)).cpu().numpy())
    assert np.max(np.abs(out - dev)) < 1e-14


def test_thread_per_circuit_kernel_against_oracle_and_warp_kernels():
    """dmx_frame16t_kernel (one thread carries one circuit: 16 touched coefficients in registers, XOR-paired per segment):
    the BASELINE sweep and ragged batches against the C oracle at 1e-10 and against the warp-per-circuit kernels to rounding,
    for every noise model whose plan offers the form."""
    torch = torch_cuda()
    terms = h2_terms()
    ham = O.pauli_terms_to_matrix(4, *terms)
    noises = {"dep": ([O.depolarize(1e-3)], [O.two_qubit_depolarize(1e-3)]),
              "pauli": ([O.asymmetric_depolarize(1e-4, 1e-4, 6e-4)], [O.two_qubit_pauli_channel(1e-4, 1e-4, 6e-4)]),
              "bitflip+dep": ([O.bit_flip(1e-3), O.depolarize(1e-3)], [O.two_qubit_depolarize(1e-3)]), "none": ([], [])}
    rng = np.random.default_rng(16)
    for name, (n1, n2) in noises.items():
        thread = build_program(h2_ops(n1, n2), 4, terms)
        warp = build_program(h2_ops(n1, n2), 4, terms, frame16t=0)
        general = build_program(h2_ops(n1, n2), 4, terms, frame16t_scaled=0)      # four-instruction form with explicit weights
        assert "f16 n=16 scaled=1" in thread.dump(), name
        for B in (65536, 2048, 2049, 4099):
            a = np.linspace(-0.5, 0.5, B).reshape(-1, 1) if B == 65536 else rng.normal(0, 0.4, size=(B, 1))
            got = thread.energies(torch.tensor(a)).cpu().numpy()
            assert np.max(np.abs(got - warp.energies(torch.tensor(a)).cpu().numpy())) < 1e-14, (name, B)
            assert np.max(np.abs(got - general.energies(torch.tensor(a)).cpu().numpy())) < 1e-14, (name, B)
            for i in (0, 1, B // 3, B - 1):
                assert abs(got[i] - oracle_energy(h2_ops(n1, n2, alpha=a[i, 0]), 4, ham)) < TOL, (name, B, i)
    # below 2048 circuits the call stays on the warp-per-circuit kernels: identical results with the form switched off
    a = rng.normal(0, 0.4, size=(2047, 1))
    assert np.array_equal(thread.energies(torch.tensor(a)).cpu().numpy(), warp.energies(torch.tensor(a)).cpu().numpy())


def random_unitary(rng, d):
    m = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d))
    q, r = np.linalg.qr(m)
    return q * (np.diag(r) / np.abs(np.diag(r)))


def random_kraus(rng, d, k):
    mats = [rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d)) for _ in range(k)]
    s = sum(m.conj().T @ m for m in mats)
    w, v = np.linalg.eigh(s)
    inv_sqrt = v @ np.diag(w ** -0.5) @ v.conj().T
    return [m @ inv_sqrt for m in mats]


def random_circuit(rng, n, depth, n_params=3):
    ops = []
    for _ in range(depth):
        kind = rng.integers(0, 6)
        if n == 1 and kind in (1, 3):
            kind = 0
        if kind == 0:
            ops.append(("u", (int(rng.integers(n)),), random_unitary(rng, 2)))
        elif kind == 1:
            a, b = rng.choice(n, 2, replace=False)
            ops.append(("u", (int(a), int(b)), random_unitary(rng, 4)))
        elif kind == 2:
            ops.append(("k", (int(rng.integers(n)),), random_kraus(rng, 2, int(rng.integers(1, 4)))))
        elif kind == 3:
            a, b = rng.choice(n, 2, replace=False)
            ops.append(("k", (int(a), int(b)), random_kraus(rng, 4, int(rng.integers(1, 4)))))
        elif kind == 4:
            ops.append(("r", (int(rng.integers(n)),), (int(rng.integers(3)), f"t{int(rng.integers(n_params))}", float(rng.normal()), float(rng.normal()))))
        else:
            a, b = rng.choice(n, 2, replace=False) if n > 1 else (0, 0)
            if n > 1:
                ops.append(("u", (int(a), int(b)), O.CNOT))
    return ops


def resolve(ops, names, values):
    out = []
    rot = {0: O.rx, 1: O.ry, 2: O.rz}
    for kind, qs, pl in ops:
        if kind == "r":
            axis, name, scale, offset = pl
            out.append(("u", qs, rot[axis](scale * values[names.index(name)] + offset)))
        else:
            out.append((kind, qs, pl))
    return out


def random_hamiltonian(rng, n, terms):
    x = rng.integers(0, 1 << n, size=terms).astype(np.uint32)
    z = rng.integers(0, 1 << n, size=terms).astype(np.uint32)
    return x, z, rng.normal(size=terms)


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 6, 7])
def test_random_circuits_density_energy_probabilities(n):
    torch = torch_cuda()
    rng = np.random.default_rng(100 + n)
    ops = random_circuit(rng, n, depth=40)
    terms = random_hamiltonian(rng, n, 12)
    ham = O.pauli_terms_to_matrix(n, *terms)
    prog = build_program(ops, n, terms)
    names = prog.param_names
    B = 5
    params = rng.normal(size=(B, max(len(names), 1)))
    p_in = torch.tensor(params[:, :len(names)]) if names else None
    rho = prog.density(p_in, batch=B)
    en = prog.energies(p_in, batch=B).cpu().numpy()
    pr = prog.probabilities(p_in, batch=B)
    for b in range(B):
        ref = O.simulate(n, resolve(ops, names, params[b]))
        assert np.max(np.abs(rho[b] - ref)) < TOL
        assert abs(en[b] - O.energy_sparse(ref, ham)) < TOL
        assert np.max(np.abs(pr[b] - O.probabilities(ref))) < TOL


def qp_noisy_variant_program(noise1, noise2, terms, **options):
    """H2 clean circuit at alpha* with a variant slot after every gate (get_updated_circuit structure)."""
    gates = O.h2_gate_list(ALPHA_STAR)
    basis = O.basis_operations()
    ops = []
    for s, (kind, qs, pl) in enumerate(gates):
        noise = noise1 if len(qs) == 1 else noise2
        ops.append((kind, qs, pl))
        for kr in noise:
            ops.append(("k", qs, kr))
        ops.append(("v", qs, (basis, s)))
        for kr in noise:
            ops.append(("ck", qs, (kr, s)))
    return build_program(ops, 4, terms, **options), gates


@pytest.mark.parametrize("options", [{}, {"segments": 0}])
def test_qp_variants_match_oracle(options):
    torch = torch_cuda()
    terms = h2_terms()
    ham = O.pauli_terms_to_matrix(4, *terms)
    n1 = [O.bit_flip(1e-3), O.depolarize(1e-3)]
    n2 = [O.two_qubit_depolarize(1e-3)]
    prog, gates = qp_noisy_variant_program(n1, n2, terms, **options)
    assert prog.n_slots == 122
    rng = np.random.default_rng(20260102)
    B = 48
    variants = np.zeros((B, 122), np.uint8)
    arity = np.array([len(g[1]) for g in gates])
    for b in range(1, B):
        k = rng.integers(1, 5)
        for s in rng.choice(122, k, replace=False):
            if b < 24:   # Pauli corrections only (what the reference distribution produces)
                variants[b, s] = rng.integers(1, 4) if arity[s] == 1 else 16 * rng.integers(0, 4) + rng.integers(0, 4)
            else:        # stress: any basis op incl. rotations and projectors
                variants[b, s] = rng.integers(0, 16) if arity[s] == 1 else rng.integers(0, 256)
    got = prog.energies(None, torch.tensor(variants)).cpu().numpy()
    host = prog.energies_host(None, variants)
    assert np.max(np.abs(host - got)) < 1e-13
    for b in range(B):
        want = oracle_energy(O.variant_ops(gates, n1, n2, variants[b]), 4, ham)
        assert abs(got[b] - want) < TOL, (b, got[b], want)


def test_thread_per_row_variant_kernel_against_oracle_and_warp_kernel():
    """dmx_frame16v_kernel (batches from 2048 variant rows: one thread per row, constant-coefficient scaled segments, diagonal
    corrections from a precomputed table): rows drawn from the reference distribution, stress rows (several Pauli corrections,
    R-type / projector entries, more entries than the per-row list holds -> general path), ragged batches, every alignment
    of the table pointer - against the C oracle at 1e-10, against the warp-per-row kernel to rounding, and bit for bit
    independent of the row's position."""
    torch = torch_cuda()
    terms = h2_terms()
    n1, n2 = [O.bit_flip(1e-3), O.depolarize(1e-3)], [O.two_qubit_depolarize(1e-3)]
    prog, gates = qp_noisy_variant_program(n1, n2, terms)
    warp, _ = qp_noisy_variant_program(n1, n2, terms, frame16t=0)
    assert "f16 n=16 scaled=0 const=1" in prog.dump()
    ham = O.pauli_terms_to_matrix(4, *terms)
    arity = np.array([len(g[1]) for g in gates])
    rng = np.random.default_rng(2718)
    B = 4099
    v = np.zeros((B, 122), np.uint8)
    for b in range(B):
        kind = b % 8
        if kind < 5:                                             # the reference distribution: mostly identity rows
            k = int(rng.random() < 0.35)
        else:
            k = int(rng.integers(1, 7))
        for s in rng.choice(122, k, replace=False):
            if kind == 7 and rng.random() < 0.3:                 # any basis operation incl. rotations and projectors
                v[b, s] = rng.integers(0, 16) if arity[s] == 1 else rng.integers(0, 256)
            else:
                v[b, s] = rng.integers(1, 4) if arity[s] == 1 else 16 * rng.integers(0, 4) + rng.integers(0, 4)
    v[7, [0, 121]] = [3, 2]                                      # first and last slot of a row
    v[11, 5] = 7                                                 # R-type correction: general path
    v[13, :40] = np.where(arity[:40] == 1, 1, 17)                # more entries than the per-row list holds: general path
    v[4098, 60] = 2                                              # last row of a ragged CTA
    base = prog.energies(None, torch.tensor(v, device="cuda")).cpu().numpy()
    ref = warp.energies(None, torch.tensor(v, device="cuda")).cpu().numpy()
    assert np.max(np.abs(base - ref)) < 1e-13
    for b in list(range(0, 64)) + [4098]:
        want = oracle_energy(O.variant_ops(gates, n1, n2, v[b]), 4, ham)
        assert abs(base[b] - want) < TOL, (b, base[b], want)
    for shift in (1, 2, 3, 31, 127):
        shifted = np.concatenate([np.zeros((shift, 122), np.uint8), v])
        got = prog.energies(None, torch.tensor(shifted, device="cuda")).cpu().numpy()
        assert np.array_equal(got[shift:], base), shift
        assert np.all(got[:shift] == got[0])
    for off in (1, 2, 3):                                        # table pointer not 4-byte aligned
        buf = torch.zeros(B * 122 + off, dtype=torch.uint8, device="cuda")
        view = buf[off:].view(B, 122)
        view.copy_(torch.tensor(v))
        assert np.array_equal(prog.energies(None, view).cpu().numpy(), base), off
    host = prog.energies_host(None, v)                           # DMA pipeline -> the same kernel
    assert np.array_equal(host, base)


def test_variant_rows_position_independent_and_unaligned_table():
    """The frame kernel walks a group's identifier rows as one run of 32-bit words: the energies of a row must not depend on
    where it sits in the batch (shifted by 1..9 identity rows: every alignment of the 122-byte rows inside the 128-byte
    steps, ragged last groups) nor on the alignment of the table pointer (byte path), bit for bit - and rows with many
    corrections or R-type entries still take the general path."""
    torch = torch_cuda()
    terms = h2_terms()
    n1, n2 = [O.bit_flip(1e-3), O.depolarize(1e-3)], [O.two_qubit_depolarize(1e-3)]
    prog, gates = qp_noisy_variant_program(n1, n2, terms)
    ham = O.pauli_terms_to_matrix(4, *terms)
    arity = np.array([len(g[1]) for g in gates])
    rng = np.random.default_rng(314)
    B = 61
    v = np.zeros((B, 122), np.uint8)
    for b in range(B):
        for s in rng.choice(122, int(rng.integers(0, 7)), replace=False):
            v[b, s] = rng.integers(1, 4) if arity[s] == 1 else 16 * rng.integers(0, 4) + rng.integers(0, 4)
    v[7, [0, 121]] = [3, 2]                                  # first and last slot of a row
    v[11, 5] = 7                                             # R-type correction: general path
    v[13, :40] = np.where(arity[:40] == 1, 1, 17)            # more entries than the compact list holds: general path
    base = prog.energies(None, torch.tensor(v, device="cuda")).cpu().numpy()
    for b in (0, 7, 11, 13, 60):
        want = oracle_energy(O.variant_ops(gates, n1, n2, v[b]), 4, ham)
        assert abs(base[b] - want) < TOL, (b, base[b], want)
    for shift in range(1, 10):
        shifted = np.concatenate([np.zeros((shift, 122), np.uint8), v])
        got = prog.energies(None, torch.tensor(shifted, device="cuda")).cpu().numpy()
        assert np.array_equal(got[shift:], base), shift
        assert np.all(got[:shift] == got[0])
    for off in (1, 2, 3):                                    # table pointer not 4-byte aligned
        buf = torch.zeros(B * 122 + off, dtype=torch.uint8, device="cuda")
        view = buf[off:].view(B, 122)
        view.copy_(torch.tensor(v, device="cuda"))
        assert view.data_ptr() % 4 == off
        assert np.array_equal(prog.energies(None, view).cpu().numpy(), base), off


def test_one_plan_on_two_streams_concurrently():
    """include/dmx.h: a finalized plan may be used from several streams at once (per-call worklist / reduction
    scratch).  Two different variant tables - both with rows that take the general path - are evaluated and reduced on
    two streams, interleaved, and must equal the results of running them one after the other."""
    torch = torch_cuda()
    from vqe_errormitigation_b200 import engine
    terms = h2_terms()
    n1, n2 = [O.depolarize(1e-3)], [O.two_qubit_depolarize(1e-3)]
    prog, gates = qp_noisy_variant_program(n1, n2, terms)
    arity = np.array([len(g[1]) for g in gates])
    rng = np.random.default_rng(77)
    B = 20000
    tables = []
    for t in range(2):
        v = np.zeros((B, 122), np.uint8)
        hot = rng.choice(B, B // 5, replace=False)
        slot = rng.integers(0, 122, hot.size)
        v[hot, slot] = np.where(arity[slot] == 1, rng.integers(1, 16, hot.size), rng.integers(1, 256, hot.size))   # incl. R / projector ops
        tables.append(torch.tensor(v, device="cuda"))
    weights = torch.ones(B, dtype=torch.float64, device="cuda")
    want = [prog.energies(None, tb).clone() for tb in tables]
    want_red = [engine.qp_reduce(w, weights).clone() for w in want]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    outs = [[], []]
    for _ in range(6):
        for i, st in enumerate(streams):
            with torch.cuda.stream(st):
                e = prog.energies(None, tables[i])
                outs[i].append((e, engine.qp_reduce(e, weights)))
    torch.cuda.synchronize()
    for i in range(2):
        for e, red in outs[i]:
            assert torch.equal(e, want[i]) and torch.equal(red, want_red[i])


@pytest.mark.parametrize("n,tile,remap", [(8, 7, 1), (8, 7, 0), (8, 5, 1), (9, 6, 1), (9, 6, 0), (9, 4, 1), (10, 6, 1)])
def test_streaming_matches_oracle(n, tile, remap):
    """Streaming passes (state in HBM) against the numpy oracle: dmx_stream_kernel with the remapping scheduler (two state
    buffers, per-pass qubit order) and with pinned low qubits (in place), the generic tile kernel for tiles below 6 qubits;
    dense two-qubit and non-unital (amplitude damping) micro-ops ride along."""
    torch = torch_cuda()
    rng = np.random.default_rng(300 + n + tile)
    ops = []
    for layer in range(3):
        for q in range(n):
            ops.append(("r", (q,), (1, f"y{layer}_{q}", 1.0, 0.0)))
            ops.append(("k", (q,), O.depolarize(1e-3)))
            ops.append(("r", (q,), (2, f"z{layer}_{q}", 1.0, 0.0)))
            ops.append(("k", (q,), O.depolarize(1e-3)))
        for q in range(n - 1):
            ops.append(("u", (q, q + 1), O.CNOT))
            ops.append(("k", (q, q + 1), O.two_qubit_depolarize(1e-3)))
    ops.append(("u", (0, n - 1), random_unitary(rng, 4)))
    ops.append(("k", (1,), O.amplitude_damp(0.01)))
    terms = random_hamiltonian(rng, n, 30)
    ham = O.pauli_terms_to_matrix(n, *terms)
    prog = build_program(ops, n, terms, tile_qubits=tile, remap=remap)
    assert prog.info["path"] == 2 and prog.info["n_passes"] >= 2
    assert prog.info["workspace_bytes_per_state"] == prog.info["state_bytes"] * (2 if remap and tile >= 6 else 1)
    B = 3
    params = rng.uniform(-np.pi, np.pi, size=(B, prog.n_params))
    en = prog.energies(torch.tensor(params), workspace_states=2).cpu().numpy()
    rho = prog.density(torch.tensor(params[:1]))
    for b in range(B):
        ref = O.simulate(n, resolve(ops, prog.param_names, params[b]))
        assert abs(en[b] - O.energy_sparse(ref, ham)) < TOL
        if b == 0:
            assert np.max(np.abs(rho[0] - ref)) < TOL


def test_qp_reduce():
    torch = torch_cuda()
    from vqe_errormitigation_b200 import engine
    rng = np.random.default_rng(9)
    for B in (1, 1000, 300001):
        v = rng.normal(size=B)
        w = rng.normal(size=B)
        c = rng.integers(1, 5, size=B)
        out = engine.qp_reduce(torch.tensor(v, device="cuda"), torch.tensor(w, device="cuda"), torch.tensor(c, device="cuda")).cpu().numpy()
        assert abs(out[0] - np.sum(w * v)) < 1e-9 * max(1, B ** 0.5)
        assert abs(out[1] - np.sum((w * v) ** 2)) < 1e-9 * B
        assert out[2] == c.sum()


def test_golden_fixture_energies():
    """Committed golden vectors (tests/golden/h2_noisy_energies.json, made by tools/make_golden_energies.py)."""
    import json
    import os
    from workloads import GOLDEN
    torch = torch_cuda()
    with open(os.path.join(GOLDEN, "h2_noisy_energies.json")) as fh:
        gold = json.load(fh)
    noises = {"depolarize1e-3+2qdepolarize1e-3": ([O.depolarize(1e-3)], [O.two_qubit_depolarize(1e-3)]),
              "depolarize1e-3_1q_only": ([O.depolarize(1e-3)], []),
              "pauli(1e-4,1e-4,6e-4)_1q+2q": ([O.asymmetric_depolarize(1e-4, 1e-4, 6e-4)], [O.two_qubit_pauli_channel(1e-4, 1e-4, 6e-4)]),
              "bit_flip1e-3_1q_only": ([O.bit_flip(1e-3)], []), "noiseless": ([], [])}
    alphas = torch.tensor(gold["alphas"], dtype=torch.float64).reshape(-1, 1)
    for name, (n1, n2) in noises.items():
        prog = build_program(h2_ops(n1, n2), 4, h2_terms())
        got = prog.energies(alphas).cpu().numpy()
        assert np.max(np.abs(got - np.array(gold["energies"][name]))) < TOL, name
    n1, n2 = [O.bit_flip(1e-3), O.depolarize(1e-3)], [O.two_qubit_depolarize(1e-3)]
    prog, _ = qp_noisy_variant_program(n1, n2, h2_terms())
    assert prog.info["path"] == 1
    rows = np.array([c["identifier"] for c in gold["variant_cases"]], np.uint8)
    got = prog.energies(None, torch.tensor(rows)).cpu().numpy()
    assert np.max(np.abs(got - np.array([c["energy"] for c in gold["variant_cases"]]))) < TOL


def _pauli_exponential_ops(n, string_qubits, paulis, sign, name, noise1, noise2):
    """One UCCSD-style Pauli-exponential block (i_wave_function.py:176-189): basis change, CNOT ladder,
    rz(2*sign*theta) on the last qubit, reverse ladder, inverse basis change — noise after every gate."""
    ops = []

    def gate(qs, m):
        ops.append(("u", qs, m))
        for kr in (noise1 if len(qs) == 1 else noise2):
            ops.append(("k", qs, kr))

    for q, p in zip(string_qubits, paulis):
        if p != "Z":
            gate((q,), O.H if p == "X" else O.rx(-np.pi / 2))
    for a, b in zip(string_qubits[:-1], string_qubits[1:]):
        gate((a, b), O.CNOT)
    ops.append(("r", (string_qubits[-1],), (2, name, 2.0 * sign, 0.0)))
    for kr in noise1:
        ops.append(("k", (string_qubits[-1],), kr))
    for a, b in reversed(list(zip(string_qubits[:-1], string_qubits[1:]))):
        gate((a, b), O.CNOT)
    for q, p in zip(string_qubits, paulis):
        if p != "Z":
            gate((q,), O.H if p == "X" else O.rx(np.pi / 2))
    return ops


def _large_circuit(n, rng, n_blocks=3):
    n1, n2 = [O.depolarize(1e-3)], [O.two_qubit_depolarize(1e-3)]
    ops = []
    for q in range(4):                                     # HF-like initial state on the first four qubits
        ops += [("u", (q,), O.rx(np.pi)), ("k", (q,), n1[0])]
    for q in range(n):
        ops += [("r", (q,), (1, f"y{q}", 1.0, 0.0)), ("k", (q,), n1[0]), ("r", (q,), (2, f"z{q}", 1.0, 0.0)), ("k", (q,), n1[0])]
    for q in range(n - 1):
        ops += [("u", (q, q + 1), O.CNOT), ("k", (q, q + 1), n2[0])]
    for blk in range(n_blocks):                            # double-excitation style strings of weight 4 .. 8
        wgt = int(rng.integers(4, 9))
        qs = sorted(int(v) for v in rng.choice(n, wgt, replace=False))
        paulis = [str(rng.choice(["X", "Y"])) for _ in qs]
        ops += _pauli_exponential_ops(n, qs, paulis, float(rng.choice([-1.0, 1.0])), f"a{blk}", n1, n2)
    return ops


def test_ten_qubit_register_against_c_oracle():
    """BASELINE config 4/5 shapes at reduced depth on 10 qubits: hardware-efficient layer + LiH-like Pauli-exponential
    blocks, depolarizing noise after every gate, 60 Pauli terms — streaming kernel vs the C oracle."""
    from oracle import c_oracle as CO
    torch = torch_cuda()
    n = 10
    rng = np.random.default_rng(510)
    ops = _large_circuit(n, rng)
    terms = random_hamiltonian(rng, n, 60)
    prog = build_program(ops, n, terms)
    assert prog.info["path"] == 2
    params = rng.uniform(-1.0, 1.0, size=(2, prog.n_params))
    got = prog.energies(torch.tensor(params), workspace_states=2).cpu().numpy()
    want = CO.energy_batch(n, prog.ops, prog.pool, terms, params=params, n_params=prog.n_params, threads=4)
    assert np.max(np.abs(got - want)) < TOL


def test_twelve_qubit_register():
    """12 qubits (rho would be 4096^2 complex128 = 256 MiB; the engine's state is 128 MiB): a light circuit against the
    C oracle, then the full LiH-like shape through two different tilings (7 and 5 resident qubits -> different pass
    decompositions must agree) and trace preservation."""
    from oracle import c_oracle as CO
    torch = torch_cuda()
    n = 12
    rng = np.random.default_rng(512)
    light = [("u", (0,), O.rx(np.pi)), ("k", (0,), O.depolarize(1e-2)), ("r", (11,), (1, "a", 1.0, 0.0)), ("k", (11,), O.bit_flip(2e-2)),
             ("u", (0, 11), O.CNOT), ("r", (5,), (0, "b", 1.0, 0.3)), ("u", (5, 6), O.CNOT), ("k", (6,), O.amplitude_damp(0.05)),
             ("u", (6,), O.H), ("k", (6,), O.asymmetric_depolarize(0.01, 0.02, 0.03))]
    # Pauli strings supported on the touched qubits {0, 5, 6, 11} so that the expectation is not trivially zero
    sub = random_hamiltonian(rng, 4, 40)
    spread = lambda m: sum(((int(m) >> (3 - i)) & 1) << (n - 1 - q) for i, q in enumerate((0, 5, 6, 11)))
    light_terms = (np.array([spread(m) for m in sub[0]], np.uint32), np.array([spread(m) for m in sub[1]], np.uint32), sub[2])
    terms = random_hamiltonian(rng, n, 40)
    prog = build_program(light, n, light_terms)
    params = rng.uniform(-1.0, 1.0, size=(1, prog.n_params))
    got = prog.energies(torch.tensor(params), workspace_states=1).cpu().numpy()
    want = CO.energy_batch(n, prog.ops, prog.pool, light_terms, params=params, n_params=prog.n_params, threads=1)
    assert abs(want[0]) > 1e-2 and np.max(np.abs(got - want)) < TOL
    ops = _large_circuit(n, rng, n_blocks=4)
    p7 = build_program(ops, n, terms, tile_qubits=7)
    p5 = build_program(ops, n, terms, tile_qubits=5)
    assert p7.info["n_passes"] < p5.info["n_passes"]
    params = rng.uniform(-1.0, 1.0, size=(2, p7.n_params))
    cols5 = [p7.param_names.index(nm) for nm in p5.param_names]
    e7 = p7.energies(torch.tensor(params), workspace_states=2).cpu().numpy()
    e5 = p5.energies(torch.tensor(params[:, cols5]), workspace_states=2).cpu().numpy()
    assert np.max(np.abs(e7 - e5)) < 1e-12
    ident = (np.array([0], np.uint32), np.array([0], np.uint32), np.array([1.0]))
    tr = build_program(ops, n, ident).energies(torch.tensor(params), workspace_states=2).cpu().numpy()
    assert np.max(np.abs(tr - 1.0)) < 1e-12


@pytest.mark.parametrize("tile", [6, 7])
def test_eleven_qubit_long_range_circuit_against_c_oracle(tile):
    """Random long-range CNOTs, Clifford gates, rotations, Pauli and amplitude-damping noise on 11 qubits: remapped passes
    with arbitrary (non-adjacent) tiles, triple sweeps with composed permutations, general micro-ops — vs the C oracle."""
    from oracle import c_oracle as CO
    torch = torch_cuda()
    n = 11
    rng = np.random.default_rng(1100 + tile)
    ops = []
    for _ in range(120):
        a, b = (int(v) for v in rng.choice(n, 2, replace=False))
        kind = int(rng.integers(0, 5))
        if kind == 0:
            ops += [("u", (a, b), O.CNOT), ("k", (a, b), O.two_qubit_depolarize(1e-2))]
        elif kind == 1:
            ops += [("r", (a,), (int(rng.integers(3)), f"t{int(rng.integers(6))}", 1.0, 0.0)), ("k", (a,), O.depolarize(1e-2))]
        elif kind == 2:
            ops += [("u", (a,), O.H), ("k", (a,), O.bit_flip(1e-2))]
        elif kind == 3:
            ops.append(("k", (a,), O.amplitude_damp(0.02)))
        else:
            ops.append(("u", (a, b), random_unitary(rng, 4)))
    terms = random_hamiltonian(rng, n, 50)
    prog = build_program(ops, n, terms, tile_qubits=tile)
    assert prog.info["path"] == 2 and prog.info["workspace_bytes_per_state"] == 2 * prog.info["state_bytes"]
    params = rng.uniform(-1.0, 1.0, size=(3, prog.n_params))
    got = prog.energies(torch.tensor(params)).cpu().numpy()
    want = CO.energy_batch(n, prog.ops, prog.pool, terms, params=params, n_params=prog.n_params, threads=3)
    assert np.max(np.abs(got - want)) < TOL


# ---- the BENCHMARKED shapes (bench.py workloads), against the C oracle -----------------------------------------------

def _large_golden(key):
    import json
    import os
    from workloads import GOLDEN
    path = os.path.join(GOLDEN, "large_energies.json")
    if not os.path.exists(path):
        return None
    with open(path) as fh:
        return json.load(fh).get(key)


def test_hea10_depth20_bench_workload_against_c_oracle():
    """BASELINE configs[3] exactly as bench.workload_hea10 builds it (10 qubits, depth 20, 1160 ops, 100 Pauli terms; 10
    remapped passes / 103 triple sweeps on its 7-qubit tiles, 18 / 112 on 6-qubit tiles - both checked): rows 0-1 of the
    bench inputs against the C oracle run live (one circuit per thread, ~1 min) and against the committed golden energies
    the same oracle produced (tools/make_golden_large.py)."""
    import bench
    from oracle import c_oracle as CO
    torch = torch_cuda()
    wl = bench.workload_hea10()
    prog = bench.get_prog(wl)
    assert prog.info["path"] == 2 and prog.info["n_passes"] >= 10 and prog.info["n_terms"] == 100 and len(prog.ops) == 1160
    params = wl["make_inputs"](0)[0][:2]
    got = prog.energies(torch.tensor(params), workspace_states=2).cpu().numpy()
    want = CO.energy_batch(10, prog.ops, prog.pool, wl["terms"], params=params, n_params=prog.n_params, threads=2)
    assert np.max(np.abs(got - want)) < TOL, (got, want)
    gold = _large_golden("hea10")
    assert gold is not None and np.max(np.abs(np.array(gold["energies"]) - want)) < 1e-12
    from vqe_errormitigation_b200 import engine
    passes = {}
    for tile in (6, 7):
        other = engine.CircuitProgram(wl["builder"], hamiltonian=wl["terms"], options={"tile_qubits": tile})
        passes[tile] = other.info["n_passes"]
        e_t = other.energies(torch.tensor(params), workspace_states=2).cpu().numpy()
        assert np.max(np.abs(e_t - want)) < TOL, (tile, e_t, want)
    assert passes[7] < passes[6] and prog.info["n_passes"] in passes.values()
    # the same rows inside a full-size launch (64 circuits in flight per pass) give the same numbers
    big = wl["make_inputs"](0)[0][:130]
    e_big = prog.energies(torch.tensor(big)).cpu().numpy()
    assert np.max(np.abs(e_big[:2] - got)) < 1e-13


def test_lih12_26_blocks_against_c_oracle_golden():
    """BASELINE configs[4] shape at 26 Pauli-exponential blocks (bench._lih_like(max_excitations=4): 12 qubits, 1084 ops,
    631 Pauli terms) against golden energies from the C oracle (it needs ~7 min per 12-qubit row, so the values are
    committed: tests/golden/large_energies.json, tools/make_golden_large.py), through tiles of 6 and 7 qubits."""
    import bench
    sys_path_tools()
    from make_golden_large import lih12_26_inputs
    from vqe_errormitigation_b200 import engine
    torch = torch_cuda()
    gold = _large_golden("lih12_26")
    if gold is None:
        pytest.skip("tests/golden/large_energies.json has no lih12_26 record")
    b, terms, n_blocks = bench._lih_like(12, max_excitations=4)
    assert n_blocks == 26 and len(b.ops) == gold["n_ops"]
    params = lih12_26_inputs(len(b.param_names))
    for tile in (6, 7):
        prog = engine.CircuitProgram(b, hamiltonian=terms, options={"tile_qubits": tile})
        assert prog.info["path"] == 2
        got = prog.energies(torch.tensor(params), workspace_states=2).cpu().numpy()
        assert np.max(np.abs(got - np.array(gold["energies"]))) < TOL, (tile, got, gold["energies"])


def test_lih12_full_640_blocks_two_tilings_trace_and_golden():
    """The full benchmarked LiH-like plan (640 blocks, 25 224 ops): tiles of 6 and 7 resident qubits give different pass
    decompositions (547 vs 492 passes) that must agree; Tr(rho) = 1; and, when the golden record exists (the oracle needs
    hours for this circuit), row 0 of the bench inputs matches it to 1e-10."""
    import bench
    from vqe_errormitigation_b200 import engine
    torch = torch_cuda()
    wl = bench.workload_lih12()
    b, terms = wl["builder"], wl["terms"]
    params = wl["make_inputs"](0)[0][:2]
    p6 = engine.CircuitProgram(b, hamiltonian=terms, options={"tile_qubits": 6})
    p7 = engine.CircuitProgram(b, hamiltonian=terms, options={"tile_qubits": 7})
    assert p6.info["n_passes"] > p7.info["n_passes"] > 100
    e6 = p6.energies(torch.tensor(params), workspace_states=2).cpu().numpy()
    e7 = p7.energies(torch.tensor(params), workspace_states=2).cpu().numpy()
    assert np.max(np.abs(e6 - e7)) < 1e-11, (e6, e7)
    ident = (np.array([0], np.uint32), np.array([0], np.uint32), np.array([1.0]))
    tr = engine.CircuitProgram(b, hamiltonian=ident, options={"tile_qubits": 6}).energies(torch.tensor(params[:1]), workspace_states=1).cpu().numpy()
    assert abs(tr[0] - 1.0) < 1e-11
    gold = _large_golden("lih12_full")
    if gold is not None:
        assert abs(e6[0] - gold["energies"][0]) < TOL and abs(e7[0] - gold["energies"][0]) < TOL


def sys_path_tools():
    import os
    import sys
    tools = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools")
    if tools not in sys.path:
        sys.path.insert(0, tools)


def test_plan_options_are_per_plan():
    """Execution switches are stored in the plan: turning the stream kernel off on one plan must not change
    how another - already finalised - remapped plan runs."""
    torch = torch_cuda()
    rng = np.random.default_rng(77)
    n = 8
    ops = []
    for q in range(n):
        ops += [("r", (q,), (1, f"y{q}", 1.0, 0.0)), ("k", (q,), O.depolarize(1e-3))]
    for q in range(n - 1):
        ops += [("u", (q, q + 1), O.CNOT), ("k", (q, q + 1), O.two_qubit_depolarize(1e-3))]
    terms = random_hamiltonian(rng, n, 10)
    a = build_program(ops, n, terms)                                   # remapped, stream kernel
    params = rng.uniform(-1, 1, size=(2, a.n_params))
    want = a.energies(torch.tensor(params)).cpu().numpy()
    b = build_program(ops, n, terms, stream_kernel=0, remap=0)           # generic tile kernel for every pass
    got_b = b.energies(torch.tensor(params)).cpu().numpy()
    got_a = a.energies(torch.tensor(params)).cpu().numpy()              # still runs (used to fail: "needs the stream kernel")
    assert np.max(np.abs(got_b - want)) < 1e-12 and np.array_equal(got_a, want)


def test_plan_is_bound_to_its_device():
    torch = torch_cuda()
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from vqe_errormitigation_b200._lib import DmxError
    prog, _ = h2_program()
    prog.energies(torch.zeros(1, 1, dtype=torch.float64, device="cuda:0"))
    with torch.cuda.device(1):
        with pytest.raises(DmxError):
            prog.energies(torch.zeros(1, 1, dtype=torch.float64, device="cuda:1"))


def test_async_host_call_matches_sync():
    torch = torch_cuda()
    from vqe_errormitigation_b200 import engine
    prog, _ = h2_program()
    B = 4099
    p = engine.pinned_empty((B, 1))
    p[:, 0] = np.linspace(-0.4, 0.4, B)
    want = engine.pinned_empty(B)
    prog.energies_host(p, out=want)                  # page-locked in and out: the synchronous zero-copy call
    pageable = prog.energies_host(p)                 # pageable result: DMA pipeline + the device-resident kernel choice
    assert np.max(np.abs(pageable - want)) < 1e-14   # (thread-per-circuit kernel there: equal to rounding)
    outs = [engine.pinned_empty(B) for _ in range(3)]
    for o in outs:
        o[:] = 0
        prog.energies_host(p, out=o, nowait=True)
    prog.host_sync()
    for o in outs:
        assert np.array_equal(o, want)
    # in-flight calls rotate over "host_streams" streams: distinct inputs per call, every stream count, full-size batches
    for streams in (1, 3, 8):
        prog_s, _ = h2_program(host_streams=streams)
        ins, outs = [], []
        for k in range(11):
            a = engine.pinned_empty((65536, 1))
            a[:, 0] = np.linspace(-0.5 + 0.01 * k, 0.5, 65536)
            ins.append(a)
            outs.append(engine.pinned_empty(65536))
            outs[-1][:] = 0
        for a, o in zip(ins, outs):
            prog_s.energies_host(a, out=o, nowait=True)
        prog_s.host_sync()
        check = engine.pinned_empty(65536)
        for a, o in zip(ins, outs):
            assert np.array_equal(o, prog_s.energies_host(a, out=check))
    # pageable buffers: the async entry point falls back to a synchronous call
    q = np.linspace(-0.4, 0.4, 17).reshape(-1, 1)
    assert np.allclose(prog.energies_host(q, nowait=True), prog.energies_host(q), atol=0, rtol=0)
    prog.host_sync()


def test_evaluate_variants_public_entry_point():
    """IErrorMitigationManager.evaluate_variants: unique rows + counts on the host path and a device tensor
    path give the same mitigated estimate as get_mu_effective on the same table."""
    torch = torch_cuda()
    from vqe_errormitigation_b200 import cirq_compat as cirq
    from vqe_errormitigation_b200.data_containers.helper_interfaces.i_noise_wrapper import INoiseModel, INoiseWrapper
    from vqe_errormitigation_b200.data_containers.model_hydrogen import HydrogenAnsatz
    from vqe_errormitigation_b200.error_mitigation import IErrorMitigationManager, TwoQubitDepolarizingChannel
    from vqe_errormitigation_b200.processors.processor_quantum import QPU
    w = HydrogenAnsatz()
    w.operator_parameters["alpha0"] = 0.0141
    noise = INoiseModel([cirq.depolarize(2e-3)], [TwoQubitDepolarizingChannel(2e-3)], "dep")
    n_w = INoiseWrapper(w, noise)
    clean = QPU.get_resolved_circuit(n_w.get_clean_circuit(), w.operator_parameters)
    mgr = IErrorMitigationManager(clean, noise, QPU.get_hamiltonian_objective_operator(w))
    np.random.seed(5)
    mgr.set_identifier_range(3000)
    _vals, mu = mgr.get_mu_effective(True, True)
    rows, counts = mgr._rows, mgr._counts
    values_h, mu_h = mgr.evaluate_variants(rows, None, counts)
    assert abs(mu_h - mu) < 1e-12 and values_h.shape == (len(rows),)
    values_d, mu_d = mgr.evaluate_variants(torch.tensor(rows, device="cuda"), None, torch.tensor(counts, device="cuda"))
    assert abs(mu_d - mu) < 1e-11 and np.max(np.abs(values_d.cpu().numpy() - values_h)) < 1e-13
