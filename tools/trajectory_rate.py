"""Shots per second of DensityMatrixSimulator.run on circuits with mid-circuit measurements (trajectories.TrajectoryProgram:
outcome-tree levels as variant batches on the GPU) beside the oracle's restatement of cirq's per-repetition loop
(oracle/dm_oracle.py::run_repeat, numpy complex128, one thread) on a bounded number of repetitions.
Usage: python tools/trajectory_rate.py [--shots 100000] [--oracle-reps 200]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import dm_oracle as O                                   # noqa: E402  (checker / CPU baseline only)
from vqe_errormitigation_b200 import cirq_compat as cirq            # noqa: E402
from vqe_errormitigation_b200 import trajectories                   # noqa: E402
from vqe_errormitigation_b200 import qp_qem_lib as Q                # noqa: E402
import test_trajectories as T                                      # noqa: E402


def swap7_with_projections():
    """7-qubit SWAP test, depolarizing noise, three projector basis operations drawn in the measurement dialect."""
    qubits = cirq.LineQubit.range(7)
    samp = Q.QPSamp(Q.swap_circuit(3, qubits), "depolarize", 1e-3)
    gates = [op for op in samp.circuit.all_operations() if not isinstance(op.gate, cirq.MeasurementGate)]
    row = np.zeros(len(gates), np.uint8)
    one_q = [s for s, op in enumerate(gates) if len(op.qubits) == 1]
    row[one_q[0]], row[one_q[5]], row[one_q[11]] = 12, 10, 11
    circuit, _ = samp.sampled_circuit_sim(row)
    return circuit, qubits


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shots", type=int, default=100000)
    ap.add_argument("--oracle-reps", type=int, default=200)
    args = ap.parse_args()
    import torch
    out = []
    for name, make in (("wide5 (5 qubits, 6 collapsed bits, 5 final)", T.circuit_wide_measurement), ("swap7 + 3 projective measurements", swap7_with_projections)):
        circuit, q = make()
        sim = cirq.DensityMatrixSimulator(seed=1)
        sim.run(circuit, repetitions=1000)                           # compiles the prefix plans
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = sim.run(circuit, repetitions=args.shots)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        n, ops = O.from_circuit(circuit, q)
        reps = args.oracle_reps if n <= 5 else max(args.oracle_reps // 10, 5)
        t1 = time.perf_counter()
        O.run_repeat(n, ops, reps, np.random.RandomState(0))
        dto = time.perf_counter() - t1
        tp, _ = trajectories.program_for_circuit(circuit, None)
        out.append({"circuit": name, "shots": args.shots, "seconds": dt, "shots_per_s": args.shots / dt,
                    "collapse_sites": len(tp.collapse_sites), "levels": len(tp.groups), "keys": sorted(res.measurements),
                    "oracle_reps": reps, "oracle_shots_per_s": reps / dto})
        print(json.dumps(out[-1]))
    return out


if __name__ == "__main__":
    main()
