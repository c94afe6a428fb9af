"""Generate tests/golden/h2_hamiltonians.json from the reference's molecular data.

    python tools/make_golden.py

Reads the packaged integral table (itself decoded from /root/reference/src/mol_data by tools/make_h2_table.py),
builds the Jordan-Wigner Hamiltonian with the repo's restatement of openfermion's algorithm and stores the
15 Pauli terms together with the reference's stored hf_energy / fci_energy known answers.
"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from vqe_errormitigation_b200.openfermion_compat import MolecularData, jordan_wigner, pauli_masks, _packaged  # noqa: E402

out = {}
for key, rec in sorted(_packaged()["molecules"].items()):
    if not rec.get("complete"):
        continue
    m = MolecularData(description=key, filename="/nonexistent/H2_" + key).load()
    h = jordan_wigner(m.get_molecular_hamiltonian())
    x, z, c = pauli_masks(h, 4)
    out[key] = {"xmask": x.tolist(), "zmask": z.tolist(), "coeff": c.tolist(), "hf_energy": m.hf_energy,
                "fci_energy": m.fci_energy, "nuclear_repulsion": m.nuclear_repulsion}
path = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "h2_hamiltonians.json")
with open(path, "w") as fh:
    json.dump({"source": "reference src/mol_data/H2_*.hdf5 -> JW (tools/make_golden.py)", "molecules": out}, fh, indent=0, sort_keys=True)
print(len(out), "molecules ->", path)
