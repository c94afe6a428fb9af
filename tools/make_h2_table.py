"""Decode the reference's ``src/mol_data/H2_*.hdf5`` saves into one small JSON table.

Run once in the build container (``/root/reference`` is not present on the GPU box):

    python tools/make_h2_table.py /root/reference/src/mol_data vqe_errormitigation_b200/moldata/h2_sto3g.json

The table carries only deterministic *inputs* and known answers of the hot path:
integrals, CCSD amplitudes (sparse), and the stored HF/MP2/CISD/CCSD/FCI energies. It is what
``MolecularData.load()`` falls back to when no ``mol_data/H2_<r>.hdf5`` sits under the cwd
(the reference would call psi4 there, ``src/data_containers/model_hydrogen.py:59-65``).
"""
import glob
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from vqe_errormitigation_b200.moldata.hdf5_lite import Hdf5File, Hdf5FormatError  # noqa: E402


def sparse(arr):
    arr = np.asarray(arr)
    return [[list(map(int, idx)), float(arr[tuple(idx)])] for idx in np.argwhere(arr != 0.0)]


def main(src_dir: str, out_path: str):
    table = {}
    for path in sorted(glob.glob(os.path.join(src_dir, "H2_*.hdf5"))):
        key = os.path.basename(path)[3:-5]
        f = Hdf5File(path)
        rec = {}
        for name in ("nuclear_repulsion", "hf_energy", "mp2_energy", "cisd_energy", "ccsd_energy", "fci_energy",
                     "n_qubits", "n_orbitals", "n_electrons"):
            try:
                v = f[name] if name in f else None
            except Hdf5FormatError:
                v = None
            if isinstance(v, np.ndarray) or isinstance(v, bytes):
                v = None
            rec[name] = v
        try:
            rec["one_body_integrals"] = np.asarray(f["one_body_integrals"]).tolist()
            rec["two_body_integrals"] = np.asarray(f["two_body_integrals"]).reshape(-1).tolist()
            rec["ccsd_single_amps"] = sparse(f["ccsd_single_amps"])
            rec["ccsd_double_amps"] = sparse(f["ccsd_double_amps"])
            rec["complete"] = rec["fci_energy"] is not None
        except (Hdf5FormatError, KeyError, ValueError) as exc:
            rec["complete"] = False
            rec["error"] = str(exc)
        table[key] = rec
        print(key, rec.get("hf_energy"), rec.get("fci_energy"), rec["complete"])
    with open(out_path, "w") as fh:
        json.dump({"source": "MiniSean/VQE_ErrorMitigation src/mol_data/H2_*.hdf5 (Psi4 1.3.2, STO-3G)",
                   "molecules": table}, fh, indent=0, sort_keys=True)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
