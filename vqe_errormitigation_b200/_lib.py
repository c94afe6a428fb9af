"""ctypes binding of ``libdmx.so`` (C ABI declared in ``include/dmx.h``).

The library is built in-tree by ``__graft_entry__.build()`` / ``python -m vqe_errormitigation_b200.build``.
There is no fallback: if the shared object is missing, importing this module raises, and every execution
entry point raises ``DmxError`` when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdmx.so")

DMX_OP_UNITARY, DMX_OP_KRAUS, DMX_OP_ROTATION, DMX_OP_VARIANT, DMX_OP_MEASURE = 1, 2, 3, 4, 5
DMX_OPF_IF_VARIANT_NONZERO = 1
DMX_MAX_QUBITS = 14


class DmxError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libdmx error {code}: {message}")
        self.code = code


class DmxOp(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_qubits", C.c_int32), ("q0", C.c_int32), ("q1", C.c_int32),
                ("mat", C.c_int32), ("n_mats", C.c_int32), ("param", C.c_int32), ("axis", C.c_int32),
                ("slot", C.c_int32), ("flags", C.c_uint32), ("scale", C.c_double), ("offset", C.c_double)]


class DmxPlanInfo(C.Structure):
    _fields_ = [("n_qubits", C.c_int32), ("n_params", C.c_int32), ("n_slots", C.c_int32), ("n_ops_in", C.c_int32),
                ("n_blocks", C.c_int32), ("n_segments", C.c_int32), ("n_passes", C.c_int32), ("tile_qubits", C.c_int32),
                ("path", C.c_int32), ("n_terms", C.c_int32), ("state_bytes", C.c_int64),
                ("flops_per_circuit", C.c_double), ("hbm_bytes_per_circuit", C.c_double),
                ("smem_bytes_per_circuit", C.c_double), ("canonical_flops_per_circuit", C.c_double),
                ("canonical_hbm_bytes_per_circuit", C.c_double), ("workspace_bytes_per_state", C.c_int64)]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


EXPORTS = {
    # name: (restype, argtypes)
    "dmx_version": (C.c_int, []),
    "dmx_last_error": (C.c_char_p, []),
    "dmx_launch_count": (C.c_int64, []),
    "dmx_plan_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(DmxOp), C.c_int, C.POINTER(C.c_double),
                                  C.c_size_t, C.POINTER(C.c_void_p)]),
    "dmx_plan_set_hamiltonian": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                           C.POINTER(C.c_double)]),
    "dmx_plan_set_hamiltonian_groups": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double)]),
    "dmx_plan_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "dmx_plan_finalize": (C.c_int, [C.c_void_p]),
    "dmx_plan_get_info": (C.c_int, [C.c_void_p, C.POINTER(DmxPlanInfo)]),
    "dmx_plan_dump": (C.c_int64, [C.c_void_p, C.c_char_p, C.c_int64]),
    "dmx_plan_destroy": (None, [C.c_void_p]),
    "dmx_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int64]),
    "dmx_energy_batch": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_size_t, C.c_void_p]),
    "dmx_energy_batch_grouped": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_size_t, C.c_void_p]),
    "dmx_density_batch": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_size_t, C.c_void_p]),
    "dmx_diag_batch": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                 C.c_size_t, C.c_void_p]),
    "dmx_pauli_batch": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_size_t, C.c_void_p]),
    "dmx_qp_reduce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "dmx_qp_sampler_create": (C.c_int, [C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_void_p)]),
    "dmx_qp_sampler_draw": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dmx_qp_sampled_energies": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dmx_qp_sampler_destroy": (None, [C.c_void_p]),
    "dmx_qp_sample_variants": (C.c_int, [C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_double), C.c_int, C.c_uint64, C.c_uint64,
                                         C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dmx_shot_expectations": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int64, C.c_void_p, C.POINTER(C.c_uint32), C.c_int, C.c_uint64, C.c_uint64,
                                        C.c_void_p, C.c_void_p]),
    "dmx_energy_batch_host": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dmx_energy_batch_host_grouped": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dmx_energy_batch_host_async": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dmx_plan_host_sync": (C.c_int, [C.c_void_p]),
    "dmx_measure_fp64_peak": (C.c_int, [C.POINTER(C.c_double), C.c_void_p]),
}

_lib = None


def load() -> C.CDLL:
    """Load libdmx.so (once).  Raises ``ImportError`` with build instructions when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build the CUDA extension first "
            f"(python -c 'import __graft_entry__ as g; g.build()' or python -m vqe_errormitigation_b200.build). "
            f"There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in EXPORTS.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        msg = load().dmx_last_error()
        raise DmxError(rc, msg.decode("utf-8", "replace") if msg else "")
