#!/usr/bin/env python
"""Golden energies of the BENCHMARKED streaming shapes, computed by the C oracle (oracle/dm_oracle.c, the per-op
restatement of cirq's density-matrix algorithm) on the exact circuits and parameter rows bench.py uses:

  hea10       bench.workload_hea10(): 10 qubits, depth 20, rows 0..1 of make_inputs(rank 0)          (~1 min per row)
  lih12_26    bench._lih_like(max_excitations=4): 26 Pauli-exponential blocks, 2 seeded rows          (~7 min per row)
  lih12_full  bench.workload_lih12(): all 640 blocks / 25 224 ops, row 0 of make_inputs(rank 0)       (hours)

    python tools/make_golden_large.py [hea10] [lih12_26] [lih12_full] [--threads T]

Writes tests/golden/large_energies.json incrementally (one key per case).  The oracle is the checker: nothing here is
imported by the product.  The GPU tests (tests/test_gpu_parity.py) and bench.py's `parity_max_abs_err` read the file.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden", "large_energies.json")


def _save(key, rec):
    data = {}
    if os.path.exists(OUT):
        with open(OUT) as fh:
            data = json.load(fh)
    data[key] = rec
    with open(OUT, "w") as fh:
        json.dump(data, fh, indent=1)
        fh.write("\n")


def lih12_26_inputs(n_params: int) -> np.ndarray:
    return np.random.default_rng(20260107).normal(scale=0.1, size=(2, n_params))


def main():
    import bench
    from oracle import c_oracle as CO
    argv = sys.argv[1:]
    threads = max(1, CO.max_threads() - 2)
    if "--threads" in argv:
        i = argv.index("--threads")
        threads = int(argv[i + 1])
        del argv[i:i + 2]
    args = [a for a in argv if not a.startswith("--")]
    which = args or ["hea10", "lih12_26", "lih12_full"]
    for key in which:
        t0 = time.time()
        if key == "hea10":
            wl = bench.workload_hea10()
            b, terms = wl["builder"], wl["terms"]
            params = wl["make_inputs"](0)[0][:2]
        elif key == "lih12_26":
            b, terms, n_blocks = bench._lih_like(12, max_excitations=4)
            assert n_blocks == 26
            params = lih12_26_inputs(len(b.param_names))
        elif key == "lih12_full":
            wl = bench.workload_lih12()
            b, terms = wl["builder"], wl["terms"]
            params = wl["make_inputs"](0)[0][:1]
        else:
            raise SystemExit(f"unknown case {key}")
        e = CO.energy_batch(b.n_qubits, list(b.ops), b.pool(), terms, params=params, n_params=len(b.param_names),
                            threads=threads, wide=True)
        _save(key, {"energies": [float(v) for v in e], "rows": int(len(e)), "n_ops": len(b.ops), "n_params": len(b.param_names),
                    "params_sha": __import__("hashlib").sha256(np.ascontiguousarray(params).tobytes()).hexdigest(),
                    "oracle": "oracle/dm_oracle.c orc_energy_batch_wide", "seconds": round(time.time() - t0, 1)})
        print(key, e, f"{time.time() - t0:.0f} s", flush=True)


if __name__ == "__main__":
    main()
