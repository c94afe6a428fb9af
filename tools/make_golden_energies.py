"""Generate tests/golden/h2_noisy_energies.json with the numpy oracle (oracle/dm_oracle.py).

    python tools/make_golden_energies.py

The reference publishes no golden vector for any noisy energy and cannot be run here
(cirq/openfermion absent), so these vectors come from the FP64 restatement of its algorithm; they pin the
CUDA path and guard the oracle itself against regressions.  Inputs: H2 at r=0.7414 (Hamiltonian from
tests/golden/h2_hamiltonians.json), the 122-gate ansatz of SURVEY App. A.3, noise after every gate.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import dm_oracle as O  # noqa: E402
from workloads import ALPHA_STAR, h2_ops, h2_terms  # noqa: E402

NOISES = {
    "depolarize1e-3+2qdepolarize1e-3": ([O.depolarize(1e-3)], [O.two_qubit_depolarize(1e-3)]),
    "depolarize1e-3_1q_only": ([O.depolarize(1e-3)], []),
    "pauli(1e-4,1e-4,6e-4)_1q+2q": ([O.asymmetric_depolarize(1e-4, 1e-4, 6e-4)], [O.two_qubit_pauli_channel(1e-4, 1e-4, 6e-4)]),
    "bit_flip1e-3_1q_only": ([O.bit_flip(1e-3)], []),
    "noiseless": ([], []),
}
ALPHAS = [0.0, ALPHA_STAR, -0.37, 0.5]

terms = h2_terms()
ham = O.pauli_terms_to_matrix(4, *terms)
out = {"alphas": ALPHAS, "energies": {}, "variant_cases": []}
for name, (n1, n2) in NOISES.items():
    out["energies"][name] = [O.energy_sparse(O.simulate(4, h2_ops(n1, n2, alpha=a)), ham) for a in ALPHAS]
    print(name, out["energies"][name])

# QP variant circuits (get_updated_circuit structure), bit_flip+depolarize / 2q depolarize, seeded identifiers
n1, n2 = [O.bit_flip(1e-3), O.depolarize(1e-3)], [O.two_qubit_depolarize(1e-3)]
gates = O.h2_gate_list(ALPHA_STAR)
rng = np.random.default_rng(20260102)
arity = [len(g[1]) for g in gates]
for case in range(8):
    ident = np.zeros(122, dtype=int)
    for s in rng.choice(122, rng.integers(0, 4), replace=False):
        if case < 5:
            ident[s] = rng.integers(1, 4) if arity[s] == 1 else 16 * rng.integers(0, 4) + rng.integers(0, 4)
        else:
            ident[s] = rng.integers(0, 16) if arity[s] == 1 else rng.integers(0, 256)
    e = O.energy_sparse(O.simulate(4, O.variant_ops(gates, n1, n2, ident)), ham)
    out["variant_cases"].append({"identifier": ident.tolist(), "energy": e})
    print("variant", case, int((ident != 0).sum()), e)
path = os.path.join(ROOT, "tests", "golden", "h2_noisy_energies.json")
with open(path, "w") as fh:
    json.dump(out, fh, indent=0)
print("->", path)
