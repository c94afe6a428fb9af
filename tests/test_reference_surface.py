"""Drop-in surface check against the reference tree itself (skipped where /root/reference is absent, e.g. on the GPU box):
every top-level function, class and method of the reference modules on the path exists under the same name in
its mirror, and every ``cirq.<name>`` / ``of.<name>`` the reference's live code spells resolves in the facades."""
import ast
import importlib
import os
import re

import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="reference tree not present")

MIRRORS = {
    "src/processors/processor_quantum.py": "processors.processor_quantum",
    "src/processors/processor_classic.py": "processors.processor_classic",
    "src/error_mitigation.py": "error_mitigation",
    "src/circuit_noise_extension.py": "circuit_noise_extension",
    "src/data_containers/model_hydrogen.py": "data_containers.model_hydrogen",
    "src/data_containers/helper_interfaces/i_noise_wrapper.py": "data_containers.helper_interfaces.i_noise_wrapper",
    "src/data_containers/helper_interfaces/i_wave_function.py": "data_containers.helper_interfaces.i_wave_function",
    "src/data_containers/helper_interfaces/i_parameter.py": "data_containers.helper_interfaces.i_parameter",
    "src/data_containers/helper_interfaces/i_collection.py": "data_containers.helper_interfaces.i_collection",
    "src/calculate_minimisation.py": "calculate_minimisation",
    "src/main.py": "main",
    "QP_QEM_lib.py": "qp_qem_lib",
}
# names of the reference that are deliberately absent, with the reason
ALLOWED_MISSING = {
    "src/processors/processor_quantum.py": set(),
    "QP_QEM_lib.py": set(),
}


@pytest.mark.parametrize("ref_path", sorted(MIRRORS))
def test_every_reference_name_has_a_mirror(ref_path):
    with open(os.path.join(REF, ref_path)) as fh:
        tree = ast.parse(fh.read())
    module = importlib.import_module("vqe_errormitigation_b200." + MIRRORS[ref_path])
    missing = []
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and not hasattr(module, node.name):
            missing.append(node.name)
        elif isinstance(node, ast.ClassDef):
            cls = getattr(module, node.name, None)
            if cls is None:
                missing.append(node.name)
                continue
            missing += [f"{node.name}.{sub.name}" for sub in node.body if isinstance(sub, ast.FunctionDef) and not hasattr(cls, sub.name)]
    assert set(missing) <= ALLOWED_MISSING.get(ref_path, set()), missing


def _live_sources():
    for rel in list(MIRRORS) + ["src/visual_testing.py", "src/plot_minimisation.py"]:
        with open(os.path.join(REF, rel)) as fh:
            text = fh.read()
        # drop comments and string literals: only names the interpreter would resolve
        tree = ast.parse(text)
        for node in ast.walk(tree):
            if isinstance(node, ast.Attribute):
                yield rel, node


def _dotted(node):
    parts = []
    while isinstance(node, ast.Attribute):
        parts.append(node.attr)
        node = node.value
    if isinstance(node, ast.Name):
        parts.append(node.id)
        return list(reversed(parts))
    return None


def test_every_cirq_and_openfermion_name_resolves():
    from vqe_errormitigation_b200 import cirq_compat, openfermion_compat
    roots = {"cirq": cirq_compat, "of": openfermion_compat, "openfermion": openfermion_compat}
    unresolved = set()
    for rel, node in _live_sources():
        chain = _dotted(node)
        if not chain or chain[0] not in roots:
            continue
        obj = roots[chain[0]]
        for name in chain[1:]:
            if not hasattr(obj, name):
                unresolved.add((rel, ".".join(chain)))
                break
            obj = getattr(obj, name)
            if not isinstance(obj, (type, type(cirq_compat))) and not hasattr(obj, "__dict__"):
                break        # reached a value (method result, constant): deeper attributes belong to run time
    # string annotations ('cirq.NOISE_MODEL_LIKE') are literals and not walked; nothing else may be missing
    assert not unresolved, sorted(unresolved)
