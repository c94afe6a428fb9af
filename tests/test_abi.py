"""The C-ABI library loads on a CPU-only machine, exports every symbol include/dmx.h declares, reports errors
through codes + dmx_last_error, and refuses to execute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from vqe_errormitigation_b200 import _lib, engine
from workloads import build_program, h2_ops, h2_terms
from oracle import dm_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "dmx.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dmx_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    lib = _lib.load()
    names = declared_functions()
    assert len(names) == 29
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/dmx.h but not exported by libdmx.so"
        assert name in _lib.EXPORTS, f"{name} has no ctypes prototype in _lib.py"
    assert lib.dmx_version() == 100


def test_struct_layout_matches_header():
    assert C.sizeof(_lib.DmxOp) == 56
    assert C.sizeof(_lib.DmxPlanInfo) == 10 * 4 + 8 + 5 * 8 + 8


def test_error_codes_and_messages():
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.dmx_plan_create(0, 0, 0, None, 0, None, 0, C.byref(h)) == -1
    assert lib.dmx_plan_create(40, 0, 0, None, 0, None, 0, C.byref(h)) == -2
    assert b"DMX_MAX_QUBITS" in lib.dmx_last_error()
    # malformed op table: qubit out of range
    b = engine.ProgramBuilder(2)
    b.add_unitary([0], np.eye(2))
    b.ops[0].q0 = 5
    with pytest.raises(_lib.DmxError) as exc:
        engine.CircuitProgram(b)
    assert exc.value.code == -1 and "qubit index" in str(exc.value)
    # conditional op without its VARIANT op
    b = engine.ProgramBuilder(2)
    b.add_kraus([0], O.depolarize(0.1), if_slot=0)
    b.n_slots = 1
    with pytest.raises(_lib.DmxError):
        engine.CircuitProgram(b)
    # option validation / call order
    prog = build_program(h2_ops([], []), 4, h2_terms())
    assert lib.dmx_plan_set_option(prog._handle, b"fuse", 0) == -5        # already finalized
    assert lib.dmx_plan_finalize(prog._handle) == -5


def test_no_cpu_execution_path():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    prog = build_program(h2_ops([], []), 4, h2_terms())
    with pytest.raises(RuntimeError, match="no CPU path"):
        prog.energies(np.zeros((1, 1)))
    # straight through the ABI: the runtime reports the missing device, nothing is computed on the host
    out = np.full(1, 123.0)
    rc = _lib.load().dmx_energy_batch_host(prog._handle, 1, np.zeros(1).ctypes.data_as(C.c_void_p), None, out.ctypes.data_as(C.c_void_p))
    assert rc == -3 and out[0] == 123.0


def test_qp_sampler_argument_checks():
    lib = _lib.load()
    h = C.c_void_p()
    n_choices = np.array([4, 2], np.int32)
    table = np.array([[1.2, -0.1, 0.0, -0.1], [0.5, 0.5, 0.0, 0.0]])
    nc, tb = n_choices.ctypes.data_as(C.POINTER(C.c_int32)), table.ctypes.data_as(C.POINTER(C.c_double))
    assert lib.dmx_qp_sampler_create(0, nc, tb, 4, C.byref(h)) == -1
    assert lib.dmx_qp_sampler_create(2, nc, tb, 4, None) == -1
    bad = np.array([5, 2], np.int32)                                  # more choices than the row stride
    assert lib.dmx_qp_sampler_create(2, bad.ctypes.data_as(C.POINTER(C.c_int32)), tb, 4, C.byref(h)) == -1
    assert b"n_choices" in lib.dmx_last_error()
    zero = np.zeros((2, 4))
    assert lib.dmx_qp_sampler_create(2, nc, zero.ctypes.data_as(C.POINTER(C.c_double)), 4, C.byref(h)) == -1
    assert b"sums to zero" in lib.dmx_last_error()
    assert lib.dmx_qp_sampler_draw(None, 1, 0, 1, None, None, None) == -1
    lib.dmx_qp_sampler_destroy(None)                                  # no-op
    import torch
    if not torch.cuda.is_available():                                 # valid tables, no device: reported, nothing drawn
        assert lib.dmx_qp_sampler_create(2, nc, tb, 4, C.byref(h)) == -3
        with pytest.raises(RuntimeError, match="no CPU path"):
            engine.QpSampler([table[0], table[1, :2]])


def test_plan_info_and_dump():
    prog = build_program(h2_ops([O.depolarize(1e-3)], [O.two_qubit_depolarize(1e-3)]), 4, h2_terms())
    info = prog.info
    assert info["n_qubits"] == 4 and info["n_params"] == 1 and info["n_ops_in"] == 244 and info["n_terms"] == 15
    assert info["path"] == 1 and info["n_segments"] == 8 and info["state_bytes"] == 8 * 256
    assert prog.dump().startswith("plan n=4")
