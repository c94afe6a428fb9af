"""Generate tests/golden/classic_minimisation.json and tests/golden/classic_minimisation_head.jsonpickle from the
reference's stored results (``/root/reference/src/classic_minimisation/*``, written by jsonpickle 1.4.1).

    python tools/make_golden_results.py

For every stored file: decode with ``vqe_errormitigation_b200.jsonpickle_compat``, re-encode and REQUIRE the text
to equal the reference's file byte for byte (this is what pins the codec on the reference's own writer), then
record the sha256 of the file and its numeric content.  The second fixture is the first two containers of
``H2_depolar_005`` in the stored wire format, so the byte-exact round trip is also testable where
``/root/reference`` does not exist.
"""
import glob
import hashlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from vqe_errormitigation_b200 import jsonpickle_compat as jp  # noqa: E402

SRC = "/root/reference/src/classic_minimisation"
GOLD = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")

out = {}
for path in sorted(glob.glob(SRC + "/*")):
    text = open(path).read()
    obj = jp.decode(text)
    assert jp.encode(obj, indent=4) == text, path
    rec = {"sha256": hashlib.sha256(text.encode()).hexdigest(), "bytes": len(text)}
    if isinstance(obj, list):
        rec["kind"] = "IContainer[]"
        rec["attributes"] = sorted(vars(obj[0]))
        rec["molecule_param"] = [float(c._molecular_parameter) for c in obj]
        rec["measured_eigenvalue"] = [float(c._measured_eigenvalue) for c in obj]
        rec["measured_std"] = [float(c._measured_standard_deviation) for c in obj if hasattr(c, "_measured_standard_deviation")]
        rec["fci_value"] = [None if c._fci_value is None else float(np.asarray(c._fci_value)) for c in obj]
    else:
        rec["kind"] = "ISimilarityCollection"
        rec["settings"] = obj._settings
        rec["mu_ideal"] = float(obj.mu_ideal)
        rec["mu_ideal_type"] = type(obj.mu_ideal).__name__
        for name in ("data_noisy", "data_mitigated"):
            h = getattr(obj, name)
            rec[name] = {"n": len(h._measurements), "mean": h._mean, "std": float(np.std(h._measurements)),
                         "first": [float(v) for v in h._measurements[:3]]}
    out[os.path.basename(path)] = rec
with open(os.path.join(GOLD, "classic_minimisation.json"), "w") as fh:
    json.dump({"source": "reference src/classic_minimisation/* decoded by tools/make_golden_results.py", "files": out}, fh, indent=1, sort_keys=True)

head = jp.decode(open(SRC + "/H2_depolar_005").read())[:2]
text = jp.encode(head, indent=4)
assert open(SRC + "/H2_depolar_005").read().startswith(text[:text.rindex("}") - 8].rstrip()[:2000])
with open(os.path.join(GOLD, "classic_minimisation_head.jsonpickle"), "w") as fh:
    fh.write(text)
print(len(out), "files ->", GOLD)
