"""``python tools/check_qpsamp_rows.py first_seed n_seeds`` - QP_QEM_lib dialect on the CPU double: for random small circuits the
batched variant program of ``QPSamp`` (P(M = 0) of every identifier row) against the oracle's simulation of the circuit
``QPSamp.sampled_circuit_sim(row)`` assembles op by op (QP_QEM_lib.py:758-795); rows use the 10 unitary basis operations, and -
for the projector entries - the unnormalised post-selected value through the measurement dialect's outcome tree."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import numpy as np                                                    # noqa: E402
import pytest                                                         # noqa: E402

import oracle_backend                                                 # noqa: E402
from oracle import dm_oracle as O                                     # noqa: E402
from vqe_errormitigation_b200 import cirq_compat as cirq              # noqa: E402
from vqe_errormitigation_b200 import engine, trajectories             # noqa: E402
from vqe_errormitigation_b200 import qp_qem_lib as L                  # noqa: E402


def main():
    first, count = int(sys.argv[1]), int(sys.argv[2])
    mp = pytest.MonkeyPatch()
    oracle_backend.install(mp)
    failures = 0
    for seed in range(first, first + count):
        rng = np.random.default_rng(seed)
        n = int(rng.integers(1, 4))
        q = cirq.LineQubit.range(n)
        ops = [cirq.H(x) for x in q]
        for _ in range(int(rng.integers(1, 6))):
            a = int(rng.integers(n))
            if n > 1 and rng.random() < 0.4:
                b = int(rng.choice([x for x in range(n) if x != a]))
                ops.append(cirq.CNOT(q[a], q[b]))
            else:
                ops.append([cirq.rx, cirq.ry, cirq.rz][int(rng.integers(3))](float(rng.normal()))(q[a]))
        ops.append(cirq.measure(q[0], key="M"))
        circuit = cirq.Circuit(ops)
        error = float(rng.choice([1e-3, 1e-2]))
        engine._PROGRAM_CACHE.clear()
        trajectories._CACHE.clear()
        samp = L.QPSamp(circuit, "depolarize", error)
        samp.qp_costs()
        gates = [op for op in circuit.all_operations() if not isinstance(op.gate, cirq.MeasurementGate)]
        rows = np.zeros((5, len(gates)), np.uint8)
        for r in range(1, 5):
            for s in rng.choice(len(gates), size=min(len(gates), int(rng.integers(1, 4))), replace=False):
                top = 10 if r < 4 else 16                       # the last row may hold projector entries
                rows[r, s] = int(rng.integers(0, top)) if len(gates[s].qubits) == 1 else 16 * int(rng.integers(0, top)) + int(rng.integers(0, top))
        prog = samp._variant_program()
        raw = prog.probabilities(None, rows, renormalise=False)
        for r in range(5):
            built, keys = samp.sampled_circuit_sim(rows[r])
            nq, oracle_ops = O.from_circuit(built, q)
            if keys:
                # post-selected, unnormalised: probability that every inserted measurement reads 0 AND M reads 0
                dist, _ = O.measurement_distribution(nq, oracle_ops)
                want = sum(p for bits, p in dist.items() if all(b == 0 for b in bits))
                got = float(samp._p0(prog, raw[r:r + 1])[0])
            else:
                rho = O.simulate(nq, [o for o in oracle_ops if o[0] != "m"])
                want = float(np.real(np.diagonal(rho)).reshape(2, -1)[0].sum())
                got = float(samp._p0(prog, raw[r:r + 1])[0])
            if abs(got - want) > 1e-10:
                failures += 1
                print(f"seed {seed} row {r}: got {got} want {want} keys {len(keys)}", flush=True)
    print(f"seeds {first}..{first + count - 1}: {failures} failure(s)")
    return 1 if failures else 0


if __name__ == "__main__":
    sys.exit(main())
