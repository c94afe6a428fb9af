"""Host side of the engine: compiles circuits to the ``dmx_op`` table of ``include/dmx.h`` and runs
batches on the GPU through ``libdmx.so``.

This is the batched form of the reference's hot path:
``QPU.get_simulated_noisy_expectation_value`` (``src/processors/processor_quantum.py:28-39``) evaluates one
resolved circuit at a time through cirq; here one compiled program evaluates ``params[B, P]`` (optimiser
parameter sets) and ``variants[B, S]`` (QP error-mitigation circuit variants, ``src/error_mitigation.py:352-399``)
per launch.  There is no CPU execution path: without a CUDA device every run method raises.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from . import cirq_compat as cirq
from ._lib import DmxOp, DmxPlanInfo, check

_AXIS = {0: "X", 1: "Y", 2: "Z"}


_TORCH_WITH_CUDA = None


def _require_cuda():
    global _TORCH_WITH_CUDA
    if _TORCH_WITH_CUDA is not None:        # a visible device does not go away: only the positive answer is cached
        return _TORCH_WITH_CUDA
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("vqe_errormitigation_b200: no CUDA device visible — the density-matrix engine has no CPU path")
    _TORCH_WITH_CUDA = torch
    return torch


def pauli_decompose(matrix: np.ndarray, tol: float = 1e-14) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Dense 2^n x 2^n observable -> (xmask, zmask, coeff) with H = sum coeff * P (bit n-1-q <-> qubit q)."""
    matrix = np.asarray(matrix)
    dim = matrix.shape[0]
    n = int(np.log2(dim))
    t = matrix.reshape((2,) * (2 * n)).astype(complex)
    # per qubit: (row bit, col bit) -> Pauli code axis of length 4 ordered (I, Z, X, Y)
    conv = np.zeros((4, 2, 2), dtype=complex)
    conv[0] = [[0.5, 0], [0, 0.5]]
    conv[1] = [[0.5, 0], [0, -0.5]]
    conv[2] = [[0, 0.5], [0.5, 0]]
    conv[3] = [[0, 0.5j], [-0.5j, 0]]   # Tr(Y M)/2 = i (M01 - M10) / 2
    for q in range(n):
        # contract row axis q and the matching column axis (always at position n after the moves)
        t = np.moveaxis(np.tensordot(conv, t, axes=([1, 2], [q, n])), 0, q)
        n_cols_left = n - 1 - q
        # after tensordot the remaining axes are: rows (n-1 of them, q removed) + cols (q's removed);
        # moveaxis put the code axis back at position q, so rows are again at 0..n-1 and the untouched
        # column axes follow; the next qubit's column axis sits at index n again.
        del n_cols_left
    coeffs = t.reshape(-1)
    xs, zs, cs = [], [], []
    for idx in np.flatnonzero(np.abs(coeffs) > tol):
        x = z = 0
        for q in range(n):
            code = (idx >> (2 * (n - 1 - q))) & 3
            bit = 1 << (n - 1 - q)
            if code & 2:
                x |= bit
            if code & 1:
                z |= bit
        c = coeffs[idx]
        if abs(c.imag) > 1e-10:
            raise ValueError("observable is not Hermitian")
        xs.append(x)
        zs.append(z)
        cs.append(c.real)
    return np.array(xs, np.uint32), np.array(zs, np.uint32), np.array(cs, np.float64)


class ProgramBuilder:
    """Accumulates a ``dmx_op`` table + matrix pool."""

    def __init__(self, n_qubits: int):
        self.n_qubits = n_qubits
        self.ops: List[DmxOp] = []
        self._pool: List[np.ndarray] = []
        self._pool_len = 0
        self._pool_index: Dict[bytes, int] = {}
        self.param_names: List[str] = []
        self.n_slots = 0

    def _add_matrices(self, mats: Sequence[np.ndarray]) -> int:
        arr = np.ascontiguousarray(np.stack([np.asarray(m, dtype=np.complex128) for m in mats]))
        key = arr.tobytes()
        if key in self._pool_index:
            return self._pool_index[key]
        off = self._pool_len
        flat = arr.view(np.float64).reshape(-1)
        self._pool.append(flat)
        self._pool_len += flat.size
        self._pool_index[key] = off
        return off

    def param_column(self, name: str) -> int:
        if name not in self.param_names:
            self.param_names.append(name)
        return self.param_names.index(name)

    def _op(self, kind, qubits, mat=-1, n_mats=0, param=-1, axis=0, slot=-1, flags=0, scale=0.0, offset=0.0):
        q = list(qubits)
        op = DmxOp(kind, len(q), q[0], q[1] if len(q) > 1 else 0, mat, n_mats, param, axis, slot, flags, scale, offset)
        self.ops.append(op)
        return op

    def add_unitary(self, qubits, matrix, *, if_slot: int = -1):
        self._check_dim(qubits, matrix)
        flags = _lib.DMX_OPF_IF_VARIANT_NONZERO if if_slot >= 0 else 0
        self._op(_lib.DMX_OP_UNITARY, qubits, self._add_matrices([matrix]), 1, slot=if_slot, flags=flags)

    def add_kraus(self, qubits, kraus_ops, *, if_slot: int = -1):
        for k in kraus_ops:
            self._check_dim(qubits, k)
        flags = _lib.DMX_OPF_IF_VARIANT_NONZERO if if_slot >= 0 else 0
        self._op(_lib.DMX_OP_KRAUS, qubits, self._add_matrices(kraus_ops), len(kraus_ops), slot=if_slot, flags=flags)

    def add_rotation(self, qubit: int, axis: int, param: str, scale: float, offset: float = 0.0):
        self._op(_lib.DMX_OP_ROTATION, [qubit], param=self.param_column(param), axis=axis, scale=scale, offset=offset)

    def add_variant(self, qubits, basis_table: Sequence[np.ndarray], slot: Optional[int] = None) -> int:
        if len(basis_table) != 16:
            raise ValueError("variant basis table must hold 16 one-qubit operations")
        if slot is None:
            slot = self.n_slots
        self.n_slots = max(self.n_slots, slot + 1)
        self._op(_lib.DMX_OP_VARIANT, qubits, self._add_matrices(basis_table), 16, slot=slot)
        return slot

    def add_measure(self, qubits):
        self._op(_lib.DMX_OP_MEASURE, list(qubits)[:2])

    def _check_dim(self, qubits, matrix):
        d = 1 << len(list(qubits))
        if np.asarray(matrix).shape != (d, d):
            raise ValueError(f"matrix shape {np.asarray(matrix).shape} does not match {len(list(qubits))} qubit(s)")

    def pool(self) -> np.ndarray:
        return np.concatenate(self._pool) if self._pool else np.zeros(0)

    def prefix(self, n_ops: int) -> "ProgramBuilder":
        """The first ``n_ops`` ops as a builder of their own (same matrix pool, parameter columns and variant columns, so
        the rows of the full program fit every prefix): the state *at* a mid-circuit measurement (trajectories.py)."""
        out = ProgramBuilder(self.n_qubits)
        out.ops = list(self.ops[:n_ops])
        out._pool, out._pool_len, out._pool_index = self._pool, self._pool_len, self._pool_index
        out.param_names = list(self.param_names)
        out.n_slots = self.n_slots
        return out


class CircuitProgram:
    """A finalized ``dmx_plan`` plus what the Python callers need (qubit order, parameter names, measurements)."""

    def __init__(self, builder: ProgramBuilder, *, qubits=None, measurements=None, hamiltonian=None, options: Optional[dict] = None,
                 coefficient_groups: Optional[np.ndarray] = None):
        """``coefficient_groups[G, T]``: G alternative weight rows for the T Pauli strings of ``hamiltonian`` (e.g. one row per
        bond length); ``energies(..., groups=g[B])`` then lets circuit b use row g[b] (``dmx_plan_set_hamiltonian_groups``)."""
        lib = _lib.load()
        self._lib = lib
        self.n_qubits = builder.n_qubits
        self.param_names = list(builder.param_names)
        self.n_params = len(self.param_names)
        self.n_slots = builder.n_slots
        self.qubits = list(qubits) if qubits is not None else list(range(self.n_qubits))
        self.measurements = list(measurements or [])
        # kept for introspection / tests (same table the C ABI received)
        self.ops = list(builder.ops)
        self.pool = builder.pool()
        self.hamiltonian = hamiltonian
        self.coefficient_groups = None if coefficient_groups is None else np.ascontiguousarray(coefficient_groups, np.float64)
        self._handle = C.c_void_p()
        ops = (DmxOp * max(len(builder.ops), 1))(*builder.ops)
        pool = np.ascontiguousarray(builder.pool(), dtype=np.float64)
        check(lib.dmx_plan_create(self.n_qubits, self.n_params, self.n_slots, ops, len(builder.ops),
                                  pool.ctypes.data_as(C.POINTER(C.c_double)), pool.size, C.byref(self._handle)))
        self.has_hamiltonian = False
        self.n_groups = 0
        try:
            if hamiltonian is not None:
                x, z, c = hamiltonian
                x = np.ascontiguousarray(x, np.uint32)
                z = np.ascontiguousarray(z, np.uint32)
                c = np.ascontiguousarray(c, np.float64)
                check(lib.dmx_plan_set_hamiltonian(self._handle, len(c), x.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                   z.ctypes.data_as(C.POINTER(C.c_uint32)), c.ctypes.data_as(C.POINTER(C.c_double))))
                self.has_hamiltonian = len(c) > 0
                self.n_groups = 0
                if coefficient_groups is not None:
                    cg = np.ascontiguousarray(coefficient_groups, np.float64)
                    if cg.ndim != 2 or cg.shape[1] != len(c):
                        raise ValueError(f"coefficient_groups must have shape [G, {len(c)}]")
                    check(lib.dmx_plan_set_hamiltonian_groups(self._handle, cg.shape[0], cg.ctypes.data_as(C.POINTER(C.c_double))))
                    self.n_groups = cg.shape[0]
            for k, v in (options or {}).items():
                check(lib.dmx_plan_set_option(self._handle, k.encode(), int(v)))
            check(lib.dmx_plan_finalize(self._handle))
        except Exception:
            lib.dmx_plan_destroy(self._handle)
            self._handle = None
            raise
        info = DmxPlanInfo()
        check(lib.dmx_plan_get_info(self._handle, C.byref(info)))
        self.info = info.as_dict()
        self._workspace = None

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h:
            try:
                self._lib.dmx_plan_destroy(h)
            except Exception:
                pass
            self._handle = None

    def dump(self) -> str:
        need = self._lib.dmx_plan_dump(self._handle, None, 0)
        buf = C.create_string_buffer(int(need))
        self._lib.dmx_plan_dump(self._handle, buf, need)
        return buf.value.decode()

    # ---- device execution ---------------------------------------------------------------------------
    def _prep(self, params, variants, batch: Optional[int]):
        torch = _require_cuda()
        dev = torch.device("cuda", torch.cuda.current_device())
        if params is None:
            if self.n_params:
                raise ValueError(f"program has parameters {self.param_names}: pass params[B, {self.n_params}]")
            p_t = None
        else:
            p_t = torch.as_tensor(params, dtype=torch.float64, device=dev)
            if p_t.dim() == 1:
                p_t = p_t.reshape(1, -1) if self.n_params != 1 or p_t.numel() == 1 else p_t.reshape(-1, 1)
            if p_t.shape[1] != self.n_params:
                raise ValueError(f"params has {p_t.shape[1]} columns, program expects {self.n_params}")
            p_t = p_t.contiguous()
        v_t = None
        if variants is not None:
            v_t = torch.as_tensor(variants, device=dev)
            if v_t.dtype != torch.uint8:
                v_t = v_t.to(torch.uint8)
            if v_t.dim() == 1:
                v_t = v_t.reshape(1, -1)
            if v_t.shape[1] != self.n_slots:
                raise ValueError(f"variants has {v_t.shape[1]} columns, program expects {self.n_slots}")
            v_t = v_t.contiguous()
        sizes = {t.shape[0] for t in (p_t, v_t) if t is not None}
        if batch is not None:
            sizes.add(batch)
        if not sizes:
            sizes = {1}
        if len(sizes) != 1:
            raise ValueError(f"inconsistent batch sizes {sorted(sizes)}")
        return torch, dev, p_t, v_t, sizes.pop()

    def _ws(self, torch, dev, nbytes: int):
        if nbytes == 0:
            return None
        if self._workspace is None or self._workspace.numel() < nbytes or self._workspace.device != dev:
            self._workspace = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        return self._workspace

    @staticmethod
    def _ptr(t):
        return C.c_void_p(t.data_ptr()) if t is not None else None

    def energies(self, params=None, variants=None, *, batch: Optional[int] = None, out=None, workspace_states: Optional[int] = None,
                 groups=None, stream=None):
        """Re Tr(rho_b H) for every row; returns a CUDA float64 tensor [B] (async on the current stream, or on ``stream`` -
        a ``torch.cuda.Stream`` or a raw ``cudaStream_t`` value; independent batches on different streams overlap).
        ``groups[B]`` (int32): coefficient row of every circuit for programs built with ``coefficient_groups``."""
        if not self.has_hamiltonian:
            raise ValueError("program was built without a Hamiltonian")
        torch, dev, p_t, v_t, B = self._prep(params, variants, batch)
        if out is None:
            out = torch.empty(B, dtype=torch.float64, device=dev)
        if B == 0:
            return out
        nbytes = 0
        if self.info["path"] == 2:
            states = workspace_states if workspace_states else min(B, 64)
            nbytes = int(self.info["workspace_bytes_per_state"]) * states
        ws = self._ws(torch, dev, nbytes)
        if stream is None:
            stream = torch.cuda.current_stream().cuda_stream
        elif not isinstance(stream, int):
            stream = stream.cuda_stream
        if nbytes and stream != torch.cuda.current_stream().cuda_stream:
            raise ValueError("streaming plans share one workspace: run them on the current stream")
        stream = C.c_void_p(stream)
        if groups is not None:
            g_t = groups if (isinstance(groups, torch.Tensor) and groups.dtype == torch.int32 and groups.device == dev
                             and groups.is_contiguous()) else torch.as_tensor(groups, device=dev).to(torch.int32).contiguous()
            if g_t.numel() != B:
                raise ValueError(f"groups has {g_t.numel()} entries for a batch of {B}")
            check(self._lib.dmx_energy_batch_grouped(self._handle, B, self._ptr(p_t), self._ptr(v_t), self._ptr(g_t), self._ptr(out),
                                                     self._ptr(ws), nbytes, stream))
            return out
        check(self._lib.dmx_energy_batch(self._handle, B, self._ptr(p_t), self._ptr(v_t), self._ptr(out), self._ptr(ws), nbytes, stream))
        return out

    def energies_host(self, params: Optional[np.ndarray] = None, variants: Optional[np.ndarray] = None, batch: Optional[int] = None,
                      out: Optional[np.ndarray] = None, nowait: bool = False, groups: Optional[np.ndarray] = None) -> np.ndarray:
        """Host-buffer call (``dmx_energy_batch_host``): numpy in, numpy out, copies included.  Page-locked arrays
        (``engine.pinned_empty``) are copied from / into in place; pageable ones go through a pinned staging buffer.
        ``nowait=True`` (``dmx_energy_batch_host_async``): the batch is queued and ``out`` is valid after
        :meth:`host_sync` - several batches can then be in flight (pass page-locked ``params`` / ``out`` that stay alive)."""
        _require_cuda()
        p = v = None
        B = batch
        if params is not None and self.n_params:
            p = params if (type(params) is np.ndarray and params.dtype == np.float64 and params.flags.c_contiguous) \
                else np.ascontiguousarray(params, np.float64)
            if p.size % self.n_params:
                raise ValueError(f"params has {p.size} entries, not a multiple of {self.n_params}")
            rows = p.size // self.n_params
            if B is not None and B != rows:
                raise ValueError(f"inconsistent batch sizes {sorted({B, rows})}")
            B = rows
        if variants is not None:
            v = variants if (type(variants) is np.ndarray and variants.dtype == np.uint8 and variants.flags.c_contiguous) \
                else np.ascontiguousarray(variants, np.uint8)
            if self.n_slots == 0 or v.size % self.n_slots:
                raise ValueError(f"variants has {v.size} entries, not a multiple of {self.n_slots}")
            rows = v.size // self.n_slots
            if B is not None and B != rows:
                raise ValueError(f"inconsistent batch sizes {sorted({B, rows})}")
            B = rows
        if B is None:
            raise ValueError("inconsistent batch sizes []")
        if out is None:
            out = np.empty(B, np.float64)
        elif out.dtype != np.float64 or out.size != B or not out.flags.c_contiguous:
            raise ValueError("out must be a contiguous float64 array of the batch size")
        # raw addresses (ndarray.ctypes builds a helper object per access: several microseconds on a 40 us call)
        if groups is not None:
            g = groups if (type(groups) is np.ndarray and groups.dtype == np.int32 and groups.flags.c_contiguous) \
                else np.ascontiguousarray(groups, np.int32)
            if g.size != B:
                raise ValueError(f"groups has {g.size} entries for a batch of {B}")
            check(self._lib.dmx_energy_batch_host_grouped(self._handle, B, p.__array_interface__["data"][0] if p is not None else None,
                                                          v.__array_interface__["data"][0] if v is not None else None,
                                                          g.__array_interface__["data"][0], out.__array_interface__["data"][0]))
            return out
        fn = self._lib.dmx_energy_batch_host_async if nowait else self._lib.dmx_energy_batch_host
        check(fn(self._handle, B, p.__array_interface__["data"][0] if p is not None else None,
                 v.__array_interface__["data"][0] if v is not None else None, out.__array_interface__["data"][0]))
        return out

    def host_sync(self):
        """Wait for every ``energies_host(..., nowait=True)`` call queued on this program (``dmx_plan_host_sync``)."""
        check(self._lib.dmx_plan_host_sync(self._handle))

    def pauli_vectors(self, params=None, variants=None, *, batch: Optional[int] = None):
        torch, dev, p_t, v_t, B = self._prep(params, variants, batch)
        out = torch.empty((B, 4 ** self.n_qubits), dtype=torch.float64, device=dev)
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        check(self._lib.dmx_pauli_batch(self._handle, B, self._ptr(p_t), self._ptr(v_t), self._ptr(out), None, 0, stream))
        return out

    def density(self, params=None, variants=None, *, batch: Optional[int] = None) -> np.ndarray:
        """final_density_matrix for every row: numpy complex128 [B, 2^n, 2^n]."""
        torch, dev, p_t, v_t, B = self._prep(params, variants, batch)
        dim = 1 << self.n_qubits
        out = torch.empty((B, dim, dim), dtype=torch.complex128, device=dev)
        nbytes = int(self.info["state_bytes"]) * B
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        check(self._lib.dmx_density_batch(self._handle, B, self._ptr(p_t), self._ptr(v_t), self._ptr(out), self._ptr(ws), nbytes, stream))
        return out.cpu().numpy()

    def probabilities(self, params=None, variants=None, *, batch: Optional[int] = None, renormalise: bool = True, as_tensor: bool = False):
        """|diag rho| (renormalised like cirq's run()) for every row: [B, 2^n]."""
        torch, dev, p_t, v_t, B = self._prep(params, variants, batch)
        dim = 1 << self.n_qubits
        out = torch.empty((B, dim), dtype=torch.float64, device=dev)
        nbytes = int(self.info["state_bytes"]) * B
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        check(self._lib.dmx_diag_batch(self._handle, B, self._ptr(p_t), self._ptr(v_t), self._ptr(out), int(renormalise), self._ptr(ws), nbytes, stream))
        return out if as_tensor else out.cpu().numpy()


def qp_reduce(values, weights, counts=None):
    """(sum w*v, sum (w*v)^2, sum counts) on the device: the local part of np.mean at error_mitigation.py:451."""
    torch = _require_cuda()
    lib = _lib.load()
    out = torch.empty(3, dtype=torch.float64, device=values.device)
    c_ptr = None
    if counts is not None:
        counts = counts.to(torch.int32).contiguous()
        c_ptr = C.c_void_p(counts.data_ptr())
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    check(lib.dmx_qp_reduce(C.c_void_p(values.data_ptr()), C.c_void_p(weights.data_ptr()), c_ptr, values.numel(),
                            C.c_void_p(out.data_ptr()), stream))
    return out


def pinned_empty(shape, dtype=np.float64) -> np.ndarray:
    """Page-locked numpy array (a view of a pinned torch tensor): the host-buffer API copies from / into it without staging."""
    torch = _require_cuda()
    t = torch.empty(tuple(np.atleast_1d(shape)), dtype=getattr(torch, np.dtype(dtype).name), pin_memory=True)
    arr = t.numpy()
    _PINNED_KEEPALIVE[arr.__array_interface__["data"][0]] = t
    return arr


_PINNED_KEEPALIVE: Dict[int, object] = {}


def pack_qp_tables(qps: Sequence[np.ndarray]) -> Tuple[np.ndarray, np.ndarray]:
    """Row-padded table [S, max K] + valid lengths, the host-side input of dmx_qp_sample_variants (pack once, reuse)."""
    S = len(qps)
    stride = max(len(q) for q in qps)
    table = np.zeros((S, stride), np.float64)
    n_choices = np.zeros(S, np.int32)
    for s, q in enumerate(qps):
        table[s, :len(q)] = np.asarray(q, np.float64)
        n_choices[s] = len(q)
    return table, n_choices


class QpSampler:
    """Device-resident sampling tables of one circuit's quasi-probabilities (``dmx_qp_sampler_create``): built once,
    every ``draw`` is a single kernel launch."""

    def __init__(self, qps):
        _require_cuda()
        self._lib = _lib.load()
        table, n_choices = qps if isinstance(qps, tuple) else pack_qp_tables(qps)
        table = np.ascontiguousarray(table, np.float64)
        n_choices = np.ascontiguousarray(n_choices, np.int32)
        self.n_slots, stride = table.shape
        handle = C.c_void_p()
        check(self._lib.dmx_qp_sampler_create(self.n_slots, n_choices.ctypes.data_as(C.POINTER(C.c_int32)),
                                              table.ctypes.data_as(C.POINTER(C.c_double)), stride, C.byref(handle)))
        self._handle = handle

    def draw(self, batch: int, seed: int, first_sample: int = 0):
        """(variants uint8 [batch, S], signs float64 [batch]) CUDA tensors; sample b is global sample first_sample + b."""
        torch = _require_cuda()
        dev = torch.device("cuda", torch.cuda.current_device())
        variants = torch.empty((batch, self.n_slots), dtype=torch.uint8, device=dev)
        signs = torch.empty(batch, dtype=torch.float64, device=dev)
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        check(self._lib.dmx_qp_sampler_draw(self._handle, C.c_uint64(seed), C.c_uint64(first_sample), batch,
                                            C.c_void_p(variants.data_ptr()), C.c_void_p(signs.data_ptr()), stream))
        return variants, signs

    def sampled_energies(self, program: "CircuitProgram", batch: int, seed: int, first_sample: int = 0):
        """(energies float64 [batch], signs float64 [batch]) CUDA tensors of the variant rows ``draw(batch, seed, first_sample)``
        would produce, without materialising them where the plan allows (``dmx_qp_sampled_energies``: one thread draws and
        evaluates one sample) - the device loop of ``IErrorMitigationManager.get_mu_effective``.  Bit-identical to
        ``program.energies(None, draw(...)[0])``."""
        torch = _require_cuda()
        dev = torch.device("cuda", torch.cuda.current_device())
        values = torch.empty(batch, dtype=torch.float64, device=dev)
        signs = torch.empty(batch, dtype=torch.float64, device=dev)
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        check(self._lib.dmx_qp_sampled_energies(program._handle, self._handle, C.c_uint64(seed), C.c_uint64(first_sample), batch,
                                                C.c_void_p(values.data_ptr()), C.c_void_p(signs.data_ptr()), stream))
        return values, signs

    def __del__(self):
        handle, self._handle = getattr(self, "_handle", None), None
        if handle:
            try:
                self._lib.dmx_qp_sampler_destroy(handle)
            except Exception:  # noqa: BLE001 - interpreter shutdown
                pass


def qp_sample_variants(qps, batch: int, seed: int, first_sample: int = 0):
    """Draw ``batch`` QP circuit variants on the device: per gate slot an index with probability |qp|/sum|qp| and the
    sample's sign — the device twin of QuasiCircuitIdentifier._get_identifiers (reference error_mitigation.py:309-335).
    ``qps``: list of quasi-probability vectors, a packed (table, n_choices) pair, or a :class:`QpSampler` (reuse it when
    drawing repeatedly).  Returns (variants uint8 [batch, S], signs float64 [batch]) CUDA tensors."""
    sampler = qps if isinstance(qps, QpSampler) else QpSampler(qps)
    return sampler.draw(batch, seed, first_sample)


def shot_expectations(probs, shots, masks: Sequence[int], seed: int, first_row: int = 0):
    """Draw ``shots`` outcomes per row of ``probs[rows, 2^n]`` (CUDA tensor, e.g. ``CircuitProgram.probabilities(...,
    as_tensor=True)``) on the device and return ``2 P(even parity on mask) - 1`` for every mask, all masks from the same
    shots: CUDA float64 tensor [rows, len(masks)] (``dmx_shot_expectations``; the sampled half of cirq's ``run()`` plus the
    parity read-out of ``measure_observable``, reference model_hydrogen.py:103-112,132-166).  ``shots``: one int, or one
    count per row (array / tensor)."""
    torch = _require_cuda()
    lib = _lib.load()
    probs = probs.contiguous()
    rows, dim = probs.shape
    m = np.ascontiguousarray(masks, np.uint32)
    out = torch.empty((rows, len(m)), dtype=torch.float64, device=probs.device)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    per_row = None
    if not isinstance(shots, (int, np.integer)):
        per_row = torch.as_tensor(shots, device=probs.device).to(torch.int64).contiguous()
        if per_row.numel() != rows:
            raise ValueError(f"shots has {per_row.numel()} entries for {rows} rows")
        shots = 0
    check(lib.dmx_shot_expectations(C.c_void_p(probs.data_ptr()), rows, dim, int(shots), CircuitProgram._ptr(per_row),
                                    m.ctypes.data_as(C.POINTER(C.c_uint32)), len(m),
                                    C.c_uint64(int(seed)), C.c_uint64(int(first_row)), C.c_void_p(out.data_ptr()), stream))
    return out


def measure_fp64_peak() -> float:
    _require_cuda()
    lib = _lib.load()
    v = C.c_double()
    check(lib.dmx_measure_fp64_peak(C.byref(v), None))
    return v.value


def launch_count() -> int:
    return int(_lib.load().dmx_launch_count())


# ---------------------------------------------------------------------------------------------------------
# circuit -> program
# ---------------------------------------------------------------------------------------------------------


def _is_clifford_angle(exponent: float) -> bool:
    return abs(exponent * 2 - round(exponent * 2)) < 1e-12


def compile_circuit(circuit: cirq.Circuit, qubit_order: Optional[Sequence] = None, *, hamiltonian=None,
                    lift_numeric_rotations: bool = False, options: Optional[dict] = None,
                    builder_hook=None) -> Tuple["CircuitProgram", Optional[np.ndarray], tuple]:
    """Lower a circuit (possibly with sympy-parameterised rx/ry/rz) to a CircuitProgram.

    Returns (program, lifted_params, structure_key).  With ``lift_numeric_rotations`` every numeric non-Clifford
    Pauli rotation becomes a ROTATION op with its own parameter column whose value is returned in
    ``lifted_params`` — resolved circuits that differ only in angles then share one plan (the COBYLA loop of
    ``src/processors/processor_classic.py:77-92`` re-resolves the same circuit with new angles every call).
    """
    builder, qubits, measurements, lifted, key = lower_circuit(circuit, qubit_order, lift_numeric_rotations)
    if builder_hook is not None:
        builder_hook(builder)
    prog = CircuitProgram(builder, qubits=qubits, measurements=measurements, hamiltonian=hamiltonian, options=options)
    return prog, (np.array([lifted], np.float64) if lift_numeric_rotations else None), key


# one-qubit basis table of a measurement site: index 0 = not (yet) measured, 1 / 2 = collapsed onto |0> / |1>
_P0 = np.array([[1, 0], [0, 0]], complex)
_P1 = np.array([[0, 0], [0, 1]], complex)
MEASUREMENT_SITE_TABLE = [np.eye(2, dtype=complex), _P0, _P1] + [np.eye(2, dtype=complex)] * 13


def lower_circuit(circuit: cirq.Circuit, qubit_order=None, lift_numeric_rotations: bool = False, sites: Optional[list] = None,
                  dephase_measurements: bool = False):
    """Circuit -> (builder, qubits, measurements, lifted angles, structure key).

    Measurements: by default only terminal ones are accepted (cirq then evolves once and samples the diagonal).  With
    ``sites`` (a list to fill) every measured qubit becomes a VARIANT slot over ``MEASUREMENT_SITE_TABLE`` and a record
    ``(slot, qubit position, n_ops before it, measurement index, bit index)`` - the collapse points of cirq's per-shot
    simulation (``DensityMatrixSimulator._run_sweep_repeat`` / ``simulate`` [cirq-0.8.2]); ``dephase_measurements`` instead
    applies the channel {|0><0|, |1><1|} (cirq's ``ignore_measurement_results=True``)."""
    qubits = list(qubit_order) if qubit_order is not None else sorted(circuit.all_qubits())
    pos = {q: i for i, q in enumerate(qubits)}
    n = len(qubits)
    if n > _lib.DMX_MAX_QUBITS:
        raise ValueError(f"{n} qubits exceed DMX_MAX_QUBITS={_lib.DMX_MAX_QUBITS}")
    b = ProgramBuilder(max(n, 1))
    measurements = []
    measured = set()
    lifted: List[float] = []
    key: List[tuple] = []

    def emit(op):
        gate = op.gate
        idx = [pos[q] for q in op.qubits]
        if isinstance(gate, cirq.MeasurementGate):
            if sites is not None:
                for bit, q in enumerate(idx):
                    n_before = len(b.ops)
                    slot = b.add_variant([q], MEASUREMENT_SITE_TABLE)
                    sites.append((slot, q, n_before, len(measurements), bit))
            elif dephase_measurements:
                for q in idx:
                    b.add_kraus([q], [_P0, _P1])
            measurements.append((gate.key, idx, gate.invert_mask))
            measured.update(idx)
            key.append(("M", tuple(idx), gate.key))
            return
        if sites is None and not dephase_measurements and measured.intersection(idx):
            raise NotImplementedError("mid-circuit measurement in an evolve-once program: use trajectories.TrajectoryProgram "
                                      "(DensityMatrixSimulator.run / simulate do)")
        if len(idx) > 2:
            dec = getattr(op, "_decompose_", None)
            if dec is not None:
                parts = dec()
            elif hasattr(gate, "_decompose_"):          # e.g. an operation rebuilt by resolve_parameters, or TOFFOLI(...)
                parts = gate._decompose_(list(op.qubits))
            else:
                raise NotImplementedError(f"{gate} acts on {len(idx)} qubits and has no decomposition")
            for sub in cirq.flatten_op_tree(parts):
                emit(sub)
            return
        axis = getattr(gate, "_axis", None)
        if cirq.is_parameterized(gate):
            if axis is None:
                raise NotImplementedError(f"parameterised gate {gate} is not an X/Y/Z rotation")
            lin = cirq.linear_angle(gate.exponent)
            if lin is None or lin[0] is None:
                raise NotImplementedError(f"exponent {gate.exponent} is not affine in one symbol")
            name, scale, offset = lin
            b.add_rotation(idx[0], axis, name, np.pi * scale, np.pi * offset)
            key.append(("R", axis, idx[0], name, scale, offset))
            return
        if lift_numeric_rotations and axis is not None and not _is_clifford_angle(float(np.real(gate.exponent))):
            name = f"__angle{len(lifted)}"
            b.add_rotation(idx[0], axis, name, 1.0, 0.0)
            lifted.append(np.pi * float(np.real(gate.exponent)))
            key.append(("L", axis, idx[0]))
            return
        u = cirq.unitary(gate, None)
        if u is not None and getattr(gate, "_mixture_", None) is None and getattr(gate, "_channel_", None) is None:
            b.add_unitary(idx, u)
        else:
            b.add_kraus(idx, cirq.channel(gate))
        key.append((gate, tuple(idx)))

    for op in circuit.all_operations():
        emit(op)
    return b, qubits, measurements, lifted, tuple(key)


_PROGRAM_CACHE: "OrderedDict[tuple, CircuitProgram]" = OrderedDict()
_PROGRAM_CACHE_SIZE = 32


def program_for_circuit(resolved: cirq.Circuit, qubit_order=None, hamiltonian=None) -> Tuple[CircuitProgram, Optional[np.ndarray]]:
    """Plan cache used by the DensityMatrixSimulator facade and the single-evaluation QPU calls: resolved
    circuits that differ only in their rotation angles map to one compiled plan."""
    builder, qubits, measurements, lifted, key = lower_circuit(resolved, qubit_order, lift_numeric_rotations=True)
    ham_key = None
    if hamiltonian is not None:
        ham_key = tuple(np.ascontiguousarray(a).tobytes() for a in hamiltonian)
    # a plan's device tables are bound to the device of its first execution (dmx.h): one cache entry per device
    try:
        device = _require_cuda().cuda.current_device()
    except RuntimeError:       # host-only use (plan inspection in the CPU tests)
        device = -1
    full_key = (tuple(qubits), key, ham_key, device)
    prog = _PROGRAM_CACHE.get(full_key)
    if prog is None:
        prog = CircuitProgram(builder, qubits=qubits, measurements=measurements, hamiltonian=hamiltonian)
        _PROGRAM_CACHE[full_key] = prog
        while len(_PROGRAM_CACHE) > _PROGRAM_CACHE_SIZE:
            _PROGRAM_CACHE.popitem(last=False)
    else:
        _PROGRAM_CACHE.move_to_end(full_key)
    params = np.array([lifted], np.float64) if prog.n_params else None
    return prog, params
