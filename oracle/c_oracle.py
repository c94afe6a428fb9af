"""ctypes loader for oracle/libdm_oracle.so — TEST INFRASTRUCTURE ONLY (see dm_oracle.c)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "libdm_oracle.so")
_lib = None


class OrcOp(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_qubits", C.c_int32), ("q0", C.c_int32), ("q1", C.c_int32), ("mat", C.c_int32),
                ("n_mats", C.c_int32), ("param", C.c_int32), ("axis", C.c_int32), ("slot", C.c_int32), ("flags", C.c_uint32),
                ("scale", C.c_double), ("offset", C.c_double)]


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            raise ImportError(f"{_PATH} missing: run `make -C oracle` (or __graft_entry__.build())")
        _lib = C.CDLL(_PATH)
        _lib.orc_energy_batch.restype = C.c_int
        _lib.orc_energy_batch_wide.restype = C.c_int
        _lib.orc_density.restype = C.c_int
        _lib.orc_max_threads.restype = C.c_int
    return _lib


def _convert_ops(ops):
    """Accepts any sequence of ctypes structs with dmx_op's field layout."""
    arr = (OrcOp * max(len(ops), 1))()
    for i, o in enumerate(ops):
        for name, _ in OrcOp._fields_:
            setattr(arr[i], name, getattr(o, name))
    return arr


def max_threads() -> int:
    return int(load().orc_max_threads())


def energy_batch(n, ops, pool, terms, params=None, variants=None, batch=None, n_params=0, n_slots=0, threads=0,
                 wide=False) -> np.ndarray:
    """``wide``: a few large circuits, one after the other, each split over ``threads`` OpenMP threads (instead of one
    circuit per thread)."""
    lib = load()
    x, z, c = (np.ascontiguousarray(terms[0], np.uint32), np.ascontiguousarray(terms[1], np.uint32), np.ascontiguousarray(terms[2], np.float64))
    pool = np.ascontiguousarray(pool, np.float64)
    p = np.ascontiguousarray(params, np.float64) if params is not None else None
    v = np.ascontiguousarray(variants, np.uint8) if variants is not None else None
    if batch is None:
        batch = p.shape[0] if p is not None else v.shape[0]
    out = np.empty(batch, np.float64)
    arr = _convert_ops(ops)
    fn = lib.orc_energy_batch_wide if wide else lib.orc_energy_batch
    rc = fn(C.c_int(n), arr, C.c_int(len(ops)), pool.ctypes.data_as(C.c_void_p), C.c_int(len(c)),
                              x.ctypes.data_as(C.c_void_p), z.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p),
                              p.ctypes.data_as(C.c_void_p) if p is not None else None, C.c_int(n_params),
                              v.ctypes.data_as(C.c_void_p) if v is not None else None, C.c_int(n_slots), C.c_int64(batch),
                              out.ctypes.data_as(C.c_void_p), C.c_int(threads))
    if rc != 0:
        raise MemoryError("orc_energy_batch failed")
    return out


def density(n, ops, pool, params=None, variants=None) -> np.ndarray:
    lib = load()
    pool = np.ascontiguousarray(pool, np.float64)
    p = np.ascontiguousarray(params, np.float64) if params is not None else None
    v = np.ascontiguousarray(variants, np.uint8) if variants is not None else None
    out = np.empty((1 << n, 1 << n), np.complex128)
    arr = _convert_ops(ops)
    rc = lib.orc_density(C.c_int(n), arr, C.c_int(len(ops)), pool.ctypes.data_as(C.c_void_p),
                         p.ctypes.data_as(C.c_void_p) if p is not None else None,
                         v.ctypes.data_as(C.c_void_p) if v is not None else None, out.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise MemoryError("orc_density failed")
    return out
