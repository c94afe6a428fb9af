"""Circuit IR with the slice of the ``cirq==0.8.2`` API that the reference touches.

cirq is a pinned, un-vendored dependency of the reference (``requirements.txt:16``) and cannot be installed
here.  This module is *not* a simulator: it only describes circuits (qubits, gates, channels, moments,
parameter resolution) with cirq's names and conventions, so the reference-shaped host code
(``processors/``, ``error_mitigation.py``, ``QP_QEM_lib.py`` mirrors) reads like the original.  The numerical
work of ``DensityMatrixSimulator.simulate / run`` is done by the CUDA engine (``engine.py`` -> ``libdmx.so``).

Conventions restated from cirq 0.8.2 [from memory, cross-checked numerically in tests/test_cirq_compat.py]:
  * big-endian qubit order, ``sorted(circuit.all_qubits())``; GridQubit sorts by (row, col), LineQubit by x;
  * ``rx(t) = exp(-i t X/2)``, ``ry``, ``rz`` likewise; ``rz(t) = diag(e^{-it/2}, e^{+it/2})``;
  * ``XPowGate(exponent=e, global_shift=s)`` has eigenphases ``exp(i pi e (s + {0,1}))``;
  * a mixture ``[(p_k, U_k)]`` acts as Kraus operators ``sqrt(p_k) U_k``;
  * ``Circuit.append`` packs operations with the EARLIEST strategy;
  * gates compare by value (they are dict keys in ``IErrorMitigationManager.get_qp_lookup``,
    ``src/error_mitigation.py:513,525``).
"""
from __future__ import annotations

import enum
import functools
import itertools
from collections import Counter
from typing import Any, Callable, Dict, Iterable, Iterator, List, Optional, Sequence, Tuple, Union

import numpy as np
import sympy

# --------------------------------------------------------------------------------------------------
# qubits
# --------------------------------------------------------------------------------------------------


@functools.total_ordering
class Qid:
    def _comparison_key(self):
        raise NotImplementedError

    def _class_rank(self):
        return type(self).__name__

    def __eq__(self, other):
        return isinstance(other, Qid) and (self._class_rank(), self._comparison_key()) == (other._class_rank(), other._comparison_key())

    def __lt__(self, other):
        return (self._class_rank(), self._comparison_key()) < (other._class_rank(), other._comparison_key())

    def __hash__(self):
        return hash((self._class_rank(), self._comparison_key()))


class GridQubit(Qid):
    def __init__(self, row: int, col: int):
        self.row, self.col = row, col

    def _comparison_key(self):
        return (self.row, self.col)

    def is_adjacent(self, other) -> bool:
        return isinstance(other, GridQubit) and abs(self.row - other.row) + abs(self.col - other.col) == 1

    @staticmethod
    def rect(rows: int, cols: int, top: int = 0, left: int = 0) -> List["GridQubit"]:
        return [GridQubit(r, c) for r in range(top, top + rows) for c in range(left, left + cols)]

    @staticmethod
    def square(diameter: int, top: int = 0, left: int = 0) -> List["GridQubit"]:
        return GridQubit.rect(diameter, diameter, top, left)

    def __repr__(self):
        return f"cirq.GridQubit({self.row}, {self.col})"

    def __str__(self):
        return f"({self.row}, {self.col})"


class LineQubit(Qid):
    def __init__(self, x: int):
        self.x = x

    def _comparison_key(self):
        return self.x

    @staticmethod
    def range(*args) -> List["LineQubit"]:
        return [LineQubit(i) for i in range(*args)]

    def is_adjacent(self, other) -> bool:
        return isinstance(other, LineQubit) and abs(self.x - other.x) == 1

    def __repr__(self):
        return f"cirq.LineQubit({self.x})"

    def __str__(self):
        return str(self.x)


class NamedQubit(Qid):
    def __init__(self, name: str):
        self.name = name

    def _comparison_key(self):
        return self.name

    def __repr__(self):
        return f"cirq.NamedQubit({self.name!r})"

    def __str__(self):
        return self.name


# --------------------------------------------------------------------------------------------------
# parameters
# --------------------------------------------------------------------------------------------------


def is_parameterized(val) -> bool:
    if isinstance(val, sympy.Basic):
        return bool(val.free_symbols)
    fn = getattr(val, "_is_parameterized_", None)
    return bool(fn()) if fn is not None else False


class ParamResolver:
    def __init__(self, param_dict: Optional[Dict[Union[str, sympy.Symbol], float]] = None):
        if isinstance(param_dict, ParamResolver):
            param_dict = param_dict.param_dict
        self.param_dict = dict(param_dict or {})

    def value_of(self, value):
        if isinstance(value, (int, float, complex, np.number)):
            return value
        if isinstance(value, str):
            return self.param_dict.get(value, sympy.Symbol(value))
        if isinstance(value, sympy.Basic):
            subs = {}
            for sym in value.free_symbols:
                if sym in self.param_dict:
                    subs[sym] = self.param_dict[sym]
                elif sym.name in self.param_dict:
                    subs[sym] = self.param_dict[sym.name]
            out = value.subs(subs)
            if not out.free_symbols:
                c = complex(out)
                return c.real if abs(c.imag) == 0 else c
            return out
        return value

    def __repr__(self):
        return f"cirq.ParamResolver({self.param_dict!r})"


def linear_angle(expr) -> Optional[Tuple[Optional[str], float, float]]:
    """Decompose ``expr`` as ``scale * symbol + offset``.  Returns (symbol name | None, scale, offset), or
    None when the expression is not affine in a single symbol."""
    if not isinstance(expr, sympy.Basic):
        return (None, 0.0, float(np.real(expr)))
    syms = list(expr.free_symbols)
    if not syms:
        return (None, 0.0, float(complex(expr).real))
    if len(syms) != 1:
        return None
    s = syms[0]
    scale = sympy.diff(expr, s)
    if scale.free_symbols:
        return None
    offset = expr.subs(s, 0)
    return (s.name, float(complex(scale).real), float(complex(offset).real))


# --------------------------------------------------------------------------------------------------
# gates and operations
# --------------------------------------------------------------------------------------------------


class Gate:
    def num_qubits(self) -> int:
        raise NotImplementedError

    def _num_qubits_(self) -> int:
        return self.num_qubits()

    def on(self, *qubits: Qid) -> "GateOperation":
        if len(qubits) != self.num_qubits():
            raise ValueError(f"{self} acts on {self.num_qubits()} qubits, got {len(qubits)}")
        return GateOperation(self, qubits)

    def __call__(self, *qubits: Qid) -> "GateOperation":
        return self.on(*qubits)

    # value equality (subclasses override _value_equality_values_)
    def _value_equality_values_(self):
        return id(self)

    def __eq__(self, other):
        if not isinstance(other, Gate) or type(self) is not type(other):
            return NotImplemented if not isinstance(other, Gate) else False
        return _freeze(self._value_equality_values_()) == _freeze(other._value_equality_values_())

    def __ne__(self, other):
        r = self.__eq__(other)
        return r if r is NotImplemented else not r

    def __hash__(self):
        return hash((type(self).__name__, _freeze(self._value_equality_values_())))

    def _is_parameterized_(self):
        return False

    def _resolve_parameters_(self, resolver: ParamResolver):
        return self

    def __pow__(self, exponent):
        return NotImplemented


def _freeze(v):
    if isinstance(v, np.ndarray):
        return ("ndarray", v.shape, v.tobytes())
    if isinstance(v, (list, tuple)):
        return tuple(_freeze(x) for x in v)
    return v


class SingleQubitGate(Gate):
    def num_qubits(self) -> int:
        return 1


class TwoQubitGate(Gate):
    def num_qubits(self) -> int:
        return 2


class ThreeQubitGate(Gate):
    def num_qubits(self) -> int:
        return 3


class GateOperation:
    def __init__(self, gate: Gate, qubits: Sequence[Qid]):
        self._gate = gate
        self._qubits = tuple(qubits)

    @property
    def gate(self) -> Gate:
        return self._gate

    @property
    def qubits(self) -> Tuple[Qid, ...]:
        return self._qubits

    def with_qubits(self, *qubits):
        return GateOperation(self._gate, qubits)

    def _unitary_(self):
        fn = getattr(self._gate, "_unitary_", None)
        return fn() if fn is not None else None

    def _is_parameterized_(self):
        return is_parameterized(self._gate)

    def _resolve_parameters_(self, resolver):
        return GateOperation(self._gate._resolve_parameters_(resolver), self._qubits)

    def __eq__(self, other):
        return isinstance(other, GateOperation) and self._gate == other._gate and self._qubits == other._qubits

    def __hash__(self):
        return hash((self._gate, self._qubits))

    def __repr__(self):
        return f"{self._gate}({', '.join(map(str, self._qubits))})"


Operation = GateOperation
OP_TREE = Any

_I2 = np.eye(2, dtype=complex)
_X = np.array([[0, 1], [1, 0]], dtype=complex)
_Y = np.array([[0, -1j], [1j, 0]], dtype=complex)
_Z = np.array([[1, 0], [0, -1]], dtype=complex)
_H = np.array([[1, 1], [1, -1]], dtype=complex) / np.sqrt(2)


class _EigenGate(Gate):
    """Gate of the form exp(i pi e s) * G^e with G an involution-like base (cirq's EigenGate)."""

    _base: np.ndarray = None
    _axis: Optional[int] = None  # 0/1/2 when the gate is a Pauli power (single-qubit rotation)
    _name = "?"

    def __init__(self, *, exponent=1.0, global_shift: float = 0.0):
        self._exponent = exponent
        self._global_shift = global_shift

    @property
    def exponent(self):
        return self._exponent

    @property
    def global_shift(self):
        return self._global_shift

    def _value_equality_values_(self):
        e = self._exponent
        if isinstance(e, sympy.Basic):
            e = sympy.srepr(e)
        return (e, self._global_shift)

    def _is_parameterized_(self):
        return is_parameterized(self._exponent)

    def _resolve_parameters_(self, resolver):
        return type(self)(exponent=resolver.value_of(self._exponent), global_shift=self._global_shift)

    def __pow__(self, exponent):
        return type(self)(exponent=self._exponent * exponent, global_shift=self._global_shift)

    def _unitary_(self) -> Optional[np.ndarray]:
        if self._is_parameterized_():
            return None
        e = float(np.real(self._exponent))
        base = self._base
        dim = base.shape[0]
        if e == round(e):
            # integer power: exact matrix (no 1e-17 dust in CNOT / X / H), times the global phase
            core = base.copy() if int(round(e)) % 2 else np.eye(dim, dtype=complex)
            phase = np.exp(1j * np.pi * e * self._global_shift) if self._global_shift else 1.0
            return phase * core
        # base has eigenvalues +-1: base^e = P+ + e^{i pi e} P-
        p_plus = (np.eye(dim) + base) / 2
        p_minus = (np.eye(dim) - base) / 2
        return np.exp(1j * np.pi * e * self._global_shift) * (p_plus + np.exp(1j * np.pi * e) * p_minus)

    def num_qubits(self):
        return int(np.log2(self._base.shape[0]))

    def __str__(self):
        if self._exponent == 1:
            return self._name
        return f"{self._name}**{self._exponent}"

    __repr__ = __str__


class XPowGate(_EigenGate):
    _base, _axis, _name = _X, 0, "X"


class YPowGate(_EigenGate):
    _base, _axis, _name = _Y, 1, "Y"


class ZPowGate(_EigenGate):
    _base, _axis, _name = _Z, 2, "Z"


class HPowGate(_EigenGate):
    _base, _name = _H, "H"


_CNOT = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=complex)
_CZ = np.diag([1, 1, 1, -1]).astype(complex)
_SWAP = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=complex)


class CNotPowGate(_EigenGate):
    _base, _name = _CNOT, "CNOT"


class CZPowGate(_EigenGate):
    _base, _name = _CZ, "CZ"


class SwapPowGate(_EigenGate):
    _base, _name = _SWAP, "SWAP"


class IdentityGate(Gate):
    def __init__(self, num_qubits: int = 1):
        self._n = num_qubits

    def num_qubits(self):
        return self._n

    def _value_equality_values_(self):
        return self._n

    def _unitary_(self):
        return np.eye(1 << self._n, dtype=complex)

    def __pow__(self, exponent):
        return self

    def __str__(self):
        return "I"

    __repr__ = __str__


class MatrixGate(Gate):
    def __init__(self, matrix: np.ndarray):
        self._matrix = np.array(matrix, dtype=complex)

    def num_qubits(self):
        return int(np.log2(self._matrix.shape[0]))

    def _value_equality_values_(self):
        return self._matrix

    def _unitary_(self):
        return self._matrix

    def __str__(self):
        return "Matrix"


class MeasurementGate(Gate):
    def __init__(self, num_qubits: int = 1, key: str = "", invert_mask: Tuple[bool, ...] = ()):
        self._n = num_qubits
        self.key = key
        self.invert_mask = tuple(invert_mask)

    def num_qubits(self):
        return self._n

    def _value_equality_values_(self):
        return (self._n, self.key, self.invert_mask)

    def __str__(self):
        return f"M('{self.key}')"

    __repr__ = __str__


def measure(*qubits: Qid, key: Optional[str] = None, invert_mask: Tuple[bool, ...] = ()) -> GateOperation:
    if key is None:
        key = ",".join(str(q) for q in qubits)
    return MeasurementGate(len(qubits), key, invert_mask).on(*qubits)


class CCXPowGate(ThreeQubitGate):
    def __init__(self, exponent=1.0):
        self._exponent = exponent

    def _value_equality_values_(self):
        return self._exponent

    def _unitary_(self):
        u = np.eye(8, dtype=complex)
        u[6:, 6:] = XPowGate(exponent=self._exponent)._unitary_()
        return u

    def _decompose_(self, qubits):
        """Textbook 6-CNOT Toffoli (cirq's own sequence is not available offline; same unitary)."""
        c0, c1, t = qubits
        tdg = T ** -1
        return [H(t), CNOT(c1, t), tdg(t), CNOT(c0, t), T(t), CNOT(c1, t), tdg(t), CNOT(c0, t), T(c1), T(t),
                H(t), CNOT(c0, c1), T(c0), tdg(c1), CNOT(c0, c1)]

    def __str__(self):
        return "TOFFOLI"


class CSwapGate(ThreeQubitGate):
    def _value_equality_values_(self):
        return ()

    def _unitary_(self):
        u = np.eye(8, dtype=complex)
        u[5, 5] = u[6, 6] = 0
        u[5, 6] = u[6, 5] = 1
        return u

    def _decompose_(self, qubits):
        """cirq 0.8.2's Fredkin decomposition [cirq-0.8.2, restated]: two 24-/21-gate CNOT + T/S/H/sqrt-X
        sequences chosen by qubit adjacency - targets not adjacent -> "control between the targets" form;
        targets adjacent but the first target not next to the control -> "outside control" with the targets
        swapped; otherwise "outside control".  Noise is attached per gate (reference QP_QEM_lib.py:410-412,
        422-448; error_mitigation.py:528-572), so the noisy SWAP-test value depends on this exact sequence.
        PIN: with SingleQubit/TwoQubitPauliChannel(1e-4, 1e-4, 6e-4) the 3+3-qubit SWAP test through this
        decomposition gives <Z0> = 0.45602391; the reference's stored run
        (src/classic_minimisation/SWAP_Similarity_P0.0001_R100_C10000) has mean 0.456066 +- 0.00085 (standard
        error) - tests/test_oracle.py::test_swap_test_noisy_value_matches_reference_stored_run."""
        c, t1, t2 = qubits
        if hasattr(t1, "is_adjacent"):
            if not t1.is_adjacent(t2):
                return self._decompose_inside_control(t1, c, t2)
            if not t1.is_adjacent(c):
                return self._decompose_outside_control(c, t2, t1)
        return self._decompose_outside_control(c, t1, t2)

    @staticmethod
    def _decompose_inside_control(target1, control, target2):
        a, b, c = target1, control, target2
        tdg = T ** -1
        return [CNOT(a, b), CNOT(b, a), CNOT(c, b), H(c), T(c), CNOT(b, c), T(a), tdg(b), T(c), CNOT(a, b), CNOT(b, c),
                T(b), tdg(c), CNOT(a, b), CNOT(b, c), (X ** 0.5)(b), tdg(c), CNOT(b, a), CNOT(b, c), CNOT(a, b),
                CNOT(b, c), H(c), (S ** -1)(c), (X ** -0.5)(a)]

    @staticmethod
    def _decompose_outside_control(control, near_target, far_target):
        a, b, c = control, near_target, far_target
        tdg = T ** -1
        sweep = [CNOT(a, b), CNOT(b, c)]
        return ([CNOT(c, b), (Y ** -0.5)(c), T(a), T(b), T(c)] + sweep + [tdg(b), T(c)] + sweep + [tdg(c)] + sweep
                + [tdg(c), (X ** 0.5)(b)] + sweep + [S(c), (X ** 0.5)(b), (X ** -0.5)(c)])

    def on(self, *qubits):
        return _DecomposableOperation(self, qubits)

    def __str__(self):
        return "CSWAP"


class _DecomposableOperation(GateOperation):
    def _decompose_(self):
        return self._gate._decompose_(self._qubits)


X = XPowGate()
Y = YPowGate()
Z = ZPowGate()
H = HPowGate()
S = ZPowGate(exponent=0.5)
T = ZPowGate(exponent=0.25)
I = IdentityGate(1)
CNOT = CNotPowGate()
CX = CNOT
CZ = CZPowGate()
SWAP = SwapPowGate()
TOFFOLI = CCX = CCXPowGate()
CSWAP = FREDKIN = CSwapGate()


def rx(rads) -> XPowGate:
    return XPowGate(exponent=rads / sympy.pi if isinstance(rads, sympy.Basic) else rads / np.pi, global_shift=-0.5)


def ry(rads) -> YPowGate:
    return YPowGate(exponent=rads / sympy.pi if isinstance(rads, sympy.Basic) else rads / np.pi, global_shift=-0.5)


def rz(rads) -> ZPowGate:
    return ZPowGate(exponent=rads / sympy.pi if isinstance(rads, sympy.Basic) else rads / np.pi, global_shift=-0.5)


Rx, Ry, Rz = rx, ry, rz

# --------------------------------------------------------------------------------------------------
# noise channels
# --------------------------------------------------------------------------------------------------


def validate_probability(p: float, p_str: str) -> float:
    if p < 0:
        raise ValueError(f"{p_str} was less than 0.")
    if p > 1:
        raise ValueError(f"{p_str} was greater than 1.")
    return p


class AsymmetricDepolarizingChannel(SingleQubitGate):
    def __init__(self, p_x: float, p_y: float, p_z: float):
        self._p_x = validate_probability(p_x, "p_x")
        self._p_y = validate_probability(p_y, "p_y")
        self._p_z = validate_probability(p_z, "p_z")
        self._p_i = 1 - validate_probability(p_x + p_y + p_z, "p_x + p_y + p_z")

    def _mixture_(self):
        return ((self._p_i, _I2), (self._p_x, _X), (self._p_y, _Y), (self._p_z, _Z))

    def _value_equality_values_(self):
        return (self._p_x, self._p_y, self._p_z)

    def __str__(self):
        return f"asymmetric_depolarize(p_x={self._p_x},p_y={self._p_y},p_z={self._p_z})"

    __repr__ = __str__


class DepolarizingChannel(SingleQubitGate):
    def __init__(self, p: float):
        self._p = validate_probability(p, "p")

    @property
    def p(self):
        return self._p

    def _mixture_(self):
        return ((1 - self._p, _I2), (self._p / 3, _X), (self._p / 3, _Y), (self._p / 3, _Z))

    def _value_equality_values_(self):
        return self._p

    def __str__(self):
        return f"depolarize(p={self._p})"

    __repr__ = __str__


class BitFlipChannel(SingleQubitGate):
    def __init__(self, p: float):
        self._p = validate_probability(p, "p")

    def _mixture_(self):
        return ((1 - self._p, _I2), (self._p, _X))

    def _value_equality_values_(self):
        return self._p

    def __str__(self):
        return f"bit_flip(p={self._p})"

    __repr__ = __str__


class PhaseFlipChannel(SingleQubitGate):
    def __init__(self, p: float):
        self._p = validate_probability(p, "p")

    def _mixture_(self):
        return ((1 - self._p, _I2), (self._p, _Z))

    def _value_equality_values_(self):
        return self._p

    def __str__(self):
        return f"phase_flip(p={self._p})"

    __repr__ = __str__


class AmplitudeDampingChannel(SingleQubitGate):
    def __init__(self, gamma: float):
        self._gamma = validate_probability(gamma, "gamma")

    def _channel_(self):
        g = self._gamma
        return (np.array([[1, 0], [0, np.sqrt(1 - g)]], dtype=complex), np.array([[0, np.sqrt(g)], [0, 0]], dtype=complex))

    def _value_equality_values_(self):
        return self._gamma

    def __str__(self):
        return f"amplitude_damp(gamma={self._gamma})"

    __repr__ = __str__


class PhaseDampingChannel(SingleQubitGate):
    def __init__(self, gamma: float):
        self._gamma = validate_probability(gamma, "gamma")

    def _channel_(self):
        g = self._gamma
        return (np.array([[1, 0], [0, np.sqrt(1 - g)]], dtype=complex), np.array([[0, 0], [0, np.sqrt(g)]], dtype=complex))

    def _value_equality_values_(self):
        return self._gamma

    def __str__(self):
        return f"phase_damp(gamma={self._gamma})"

    __repr__ = __str__


def depolarize(p: float) -> DepolarizingChannel:
    return DepolarizingChannel(p)


def asymmetric_depolarize(p_x: float, p_y: float, p_z: float) -> AsymmetricDepolarizingChannel:
    return AsymmetricDepolarizingChannel(p_x, p_y, p_z)


def bit_flip(p: float) -> BitFlipChannel:
    return BitFlipChannel(p)


def phase_flip(p: float) -> PhaseFlipChannel:
    return PhaseFlipChannel(p)


def amplitude_damp(gamma: float) -> AmplitudeDampingChannel:
    return AmplitudeDampingChannel(gamma)


def phase_damp(gamma: float) -> PhaseDampingChannel:
    return PhaseDampingChannel(gamma)


# --------------------------------------------------------------------------------------------------
# protocols
# --------------------------------------------------------------------------------------------------

_RAISE = object()


def unitary(val, default=_RAISE) -> np.ndarray:
    if isinstance(val, np.ndarray):
        return val
    if isinstance(val, Circuit):
        return val.unitary()
    fn = getattr(val, "_unitary_", None)
    result = fn() if fn is not None else None
    if result is None or result is NotImplemented:
        if default is _RAISE:
            raise TypeError(f"object of type {type(val).__name__} has no unitary (parameterized gate or channel?)")
        return default
    return np.asarray(result)


def has_unitary(val) -> bool:
    return unitary(val, None) is not None


def mixture(val, default=_RAISE):
    gate = val.gate if isinstance(val, GateOperation) else val
    fn = getattr(gate, "_mixture_", None)
    if fn is not None:
        result = fn()
        if result is not NotImplemented and result is not None:
            return tuple(result)
    u = unitary(gate, None)
    if u is not None:
        return ((1.0, u),)
    if default is _RAISE:
        raise TypeError(f"object of type {type(gate).__name__} has no _mixture_ or _unitary_")
    return default


def has_mixture(val) -> bool:
    return mixture(val, None) is not None


def channel(val, default=_RAISE):
    """Kraus operators (cirq 0.8.2 calls this protocol ``channel``)."""
    gate = val.gate if isinstance(val, GateOperation) else val
    fn = getattr(gate, "_channel_", None)
    if fn is not None:
        result = fn()
        if result is not NotImplemented and result is not None:
            return tuple(np.asarray(k) for k in result)
    mix = mixture(gate, None)
    if mix is not None:
        return tuple(np.sqrt(p) * np.asarray(u) for p, u in mix)
    if default is _RAISE:
        raise TypeError(f"object of type {type(gate).__name__} has no _channel_, _mixture_ or _unitary_")
    return default


kraus = channel


def resolve_parameters(val, param_resolver):
    if not isinstance(param_resolver, ParamResolver):
        param_resolver = ParamResolver(param_resolver)
    fn = getattr(val, "_resolve_parameters_", None)
    if fn is None:
        return val
    return fn(param_resolver)


# --------------------------------------------------------------------------------------------------
# circuits
# --------------------------------------------------------------------------------------------------


class InsertStrategy(enum.Enum):
    NEW = "NEW"
    NEW_THEN_INLINE = "NEW_THEN_INLINE"
    INLINE = "INLINE"
    EARLIEST = "EARLIEST"


class Moment:
    def __init__(self, operations: Iterable[GateOperation] = ()):
        self.operations: Tuple[GateOperation, ...] = tuple(operations)

    @property
    def qubits(self):
        return frozenset(q for op in self.operations for q in op.qubits)

    def operates_on(self, qubits) -> bool:
        mine = self.qubits
        return any(q in mine for q in qubits)

    def with_operation(self, op):
        return Moment(self.operations + (op,))

    def __iter__(self):
        return iter(self.operations)

    def __len__(self):
        return len(self.operations)


def flatten_op_tree(tree) -> Iterator[GateOperation]:
    if isinstance(tree, GateOperation):
        yield tree
        return
    if isinstance(tree, Moment):
        yield from tree.operations
        return
    if isinstance(tree, (str, bytes)) or not hasattr(tree, "__iter__"):
        raise TypeError(f"not an operation tree: {tree!r}")
    for sub in tree:
        yield from flatten_op_tree(sub)


class Circuit:
    def __init__(self, *contents, strategy: InsertStrategy = InsertStrategy.EARLIEST):
        self._moments: List[Moment] = []
        for c in contents:
            self.append(c, strategy=strategy)

    @property
    def moments(self) -> List[Moment]:
        return self._moments

    def append(self, op_tree, strategy: InsertStrategy = InsertStrategy.EARLIEST):
        if isinstance(op_tree, Moment):
            self._moments.append(op_tree)
            return
        for op in flatten_op_tree(op_tree):
            if strategy in (InsertStrategy.NEW, InsertStrategy.NEW_THEN_INLINE):
                self._moments.append(Moment([op]))
                if strategy is InsertStrategy.NEW_THEN_INLINE:
                    strategy = InsertStrategy.INLINE
                continue
            if strategy is InsertStrategy.INLINE:
                if self._moments and not self._moments[-1].operates_on(op.qubits):
                    self._moments[-1] = self._moments[-1].with_operation(op)
                else:
                    self._moments.append(Moment([op]))
                continue
            # EARLIEST: slide back to just after the last moment touching any of the op's qubits
            k = len(self._moments)
            while k > 0 and not self._moments[k - 1].operates_on(op.qubits):
                k -= 1
            if k == len(self._moments):
                self._moments.append(Moment([op]))
            else:
                self._moments[k] = self._moments[k].with_operation(op)

    def all_operations(self) -> Iterator[GateOperation]:
        for m in self._moments:
            yield from m.operations

    def all_qubits(self) -> frozenset:
        return frozenset(q for m in self._moments for q in m.qubits)

    def copy(self) -> "Circuit":
        out = Circuit()
        out._moments = list(self._moments)
        return out

    def __copy__(self):
        return self.copy()

    def __iter__(self):
        return iter(self._moments)

    def __len__(self):
        return len(self._moments)

    def __add__(self, other: "Circuit") -> "Circuit":
        out = self.copy()
        out._moments.extend(other._moments)
        return out

    def has_measurements(self) -> bool:
        return any(isinstance(op.gate, MeasurementGate) for op in self.all_operations())

    def are_all_measurements_terminal(self) -> bool:
        """No operation (a later measurement included) acts on a measured qubit afterwards (cirq ``Circuit`` method that
        decides between evolve-once sampling and per-repetition simulation)."""
        measured = set()
        for op in self.all_operations():
            if measured.intersection(op.qubits):
                return False
            if isinstance(op.gate, MeasurementGate):
                measured.update(op.qubits)
        return True

    def _is_parameterized_(self):
        return any(is_parameterized(op) for op in self.all_operations())

    def _resolve_parameters_(self, resolver):
        out = Circuit()
        out._moments = [Moment(resolve_parameters(op, resolver) for op in m.operations) for m in self._moments]
        return out

    def unitary(self, qubit_order: Optional[Sequence[Qid]] = None) -> np.ndarray:
        """Dense unitary; terminal measurements are ignored (cirq behaviour used by QP_QEM_lib.py:485)."""
        qubits = list(qubit_order) if qubit_order is not None else sorted(self.all_qubits())
        n = len(qubits)
        pos = {q: i for i, q in enumerate(qubits)}
        state = np.eye(1 << n, dtype=complex).reshape((2,) * n + (1 << n,))
        for op in self.all_operations():
            if isinstance(op.gate, MeasurementGate):
                continue
            u = unitary(op)
            k = len(op.qubits)
            axes = [pos[q] for q in op.qubits]
            state = np.tensordot(u.reshape((2,) * (2 * k)), state, axes=(list(range(k, 2 * k)), axes))
            state = np.moveaxis(state, list(range(k)), axes)
        return state.reshape(1 << n, 1 << n)

    def __repr__(self):
        lines = []
        for i, m in enumerate(self._moments):
            lines.append(f"{i:4d}: " + "  ".join(repr(op) for op in m.operations))
        return "\n".join(lines)

    @staticmethod
    def _wire_symbols(op) -> Tuple[str, ...]:
        """Diagram symbols of one operation, one per qubit: the gate's own ``_circuit_diagram_info_`` when it has one (the
        reference's channel classes do, error_mitigation.py:784-786 ...), cirq's usual glyphs for the common gates otherwise."""
        gate, k = op.gate, len(op.qubits)
        info = getattr(gate, "_circuit_diagram_info_", None)
        if info is not None:
            symbols = tuple(info(CircuitDiagramInfoArgs(known_qubits=op.qubits, known_qubit_count=k)).wire_symbols)
            if len(symbols) == k:
                return symbols
        if isinstance(gate, MeasurementGate):
            return (f"M('{gate.key}')",) + ("M",) * (k - 1)
        exponent = getattr(gate, "_exponent", 1)
        power = "" if exponent == 1 else f"^{exponent:.4g}" if isinstance(exponent, (int, float)) else f"^({exponent})"
        if isinstance(gate, CNotPowGate):
            return ("@", "X" + power)
        if isinstance(gate, CZPowGate):
            return ("@", "@" + power)
        if isinstance(gate, SwapPowGate):
            return ("\u00d7", "\u00d7" + power)
        if isinstance(gate, CCXPowGate):
            return ("@", "@", "X" + power)
        if isinstance(gate, CSwapGate):
            return ("@", "\u00d7", "\u00d7")
        if isinstance(gate, (XPowGate, YPowGate, ZPowGate)) and getattr(gate, "_global_shift", 0) == -0.5:
            turns = f"{exponent:.4g}\u03c0" if isinstance(exponent, (int, float)) else f"{exponent}*\u03c0"
            return (f"R{gate._name.lower()}({turns})",)
        if isinstance(gate, _EigenGate):
            return (gate._name + power,) + tuple(f"#{i + 2}" for i in range(k - 1))
        label = str(gate)
        return (label,) + tuple(f"#{i + 2}" for i in range(k - 1))

    def to_text_diagram(self, qubit_order: Optional[Sequence[Qid]] = None) -> str:
        """cirq-style text diagram: one wire per qubit, one column per moment, vertical bars joining the wires of a
        multi-qubit operation (what ``print(circuit)`` shows in the reference's drivers, main.py:50)."""
        qubits = list(qubit_order) if qubit_order is not None else sorted(self.all_qubits())
        if not qubits:
            return ""
        row = {q: 2 * i for i, q in enumerate(qubits)}
        n_rows = 2 * len(qubits) - 1
        labels = [str(q) + ": " for q in qubits]
        width = max(len(lab) for lab in labels)
        lines = [(labels[r // 2].ljust(width) if r % 2 == 0 else " " * width) for r in range(n_rows)]
        for moment in self._moments:
            cells = [None] * n_rows
            for op in moment.operations:
                symbols = self._wire_symbols(op)
                rows = [row[q] for q in op.qubits]
                for r in range(min(rows) + 1, max(rows)):
                    cells[r] = "\u253c" if r % 2 == 0 else "\u2502"
                for r, sym in zip(rows, symbols):
                    cells[r] = sym
            col = max([len(c) for c in cells if c is not None] + [1])
            for r in range(n_rows):
                fill = "\u2500" if r % 2 == 0 else " "
                cell = cells[r] if cells[r] is not None else fill
                pad = col - len(cell)
                if r % 2 == 0:
                    lines[r] += fill * 3 + cell + fill * pad
                else:
                    lines[r] += " " * 3 + cell + " " * pad
        for r in range(0, n_rows, 2):
            lines[r] += "\u2500" * 3
        return "\n".join(line.rstrip() for line in lines)

    def __str__(self):
        return self.to_text_diagram()


# --------------------------------------------------------------------------------------------------
# noise-model base (INoiseWrapper inherits from cirq.NoiseModel, i_noise_wrapper.py:73)
# --------------------------------------------------------------------------------------------------


class NoiseModel:
    def __init__(self):
        pass

    def noisy_operation(self, op):
        return op

    def is_virtual_moment(self, moment) -> bool:
        return False


class ConstantQubitNoiseModel(NoiseModel):
    def __init__(self, qubit_noise_gate: Gate):
        super().__init__()
        self.qubit_noise_gate = qubit_noise_gate

    def noisy_operation(self, operation):
        return [operation, [self.qubit_noise_gate.on(q) for q in operation.qubits]]


# --------------------------------------------------------------------------------------------------
# simulator facade (numerics in the CUDA engine)
# --------------------------------------------------------------------------------------------------


class TrialResult:
    def __init__(self, measurements: Dict[str, np.ndarray], repetitions: int):
        self.measurements = measurements
        self.repetitions = repetitions

    def histogram(self, *, key: str, fold_func: Callable = None) -> Counter:
        bits = self.measurements[key]
        if fold_func is None:
            weights = 1 << np.arange(bits.shape[1] - 1, -1, -1)
            vals = bits.astype(np.int64) @ weights
            return Counter(int(v) for v in vals)
        return Counter(fold_func(tuple(row)) for row in bits)

    def multi_measurement_histogram(self, *, keys: Sequence[str]) -> Counter:
        cols = []
        for k in keys:
            bits = self.measurements[k]
            weights = 1 << np.arange(bits.shape[1] - 1, -1, -1)
            cols.append(bits.astype(np.int64) @ weights)
        if not cols:
            return Counter()
        stacked = np.stack(cols, axis=1)
        uniq, counts = np.unique(stacked, axis=0, return_counts=True)
        return Counter({tuple(int(v) for v in row): int(c) for row, c in zip(uniq, counts)})


class DensityMatrixTrialResult:
    def __init__(self, final_density_matrix: np.ndarray, qubits: Sequence[Qid], params: ParamResolver):
        self.final_density_matrix = final_density_matrix
        self.qubits = list(qubits)
        self.params = params


class DensityMatrixSimulator:
    """Facade with cirq's call surface; ``simulate`` / ``run`` evolve on the GPU (engine.CircuitProgram).

    Differences from cirq 0.8.2, on purpose: dtype defaults to complex128 (cirq: complex64 — SURVEY.md §0
    item 2; the north star requires FP64).  Circuits whose measurements are all terminal are evolved once and sampled from
    the diagonal; measurements with gates after them go through ``trajectories.TrajectoryProgram`` (cirq: one simulation
    per repetition with ``measure_density_matrix`` collapses; here: the outcome tree walked level by level as variant
    batches).  ``simulate`` collapses every measurement of its single run, like cirq's; ``ignore_measurement_results``
    replaces the collapse by the dephasing channel.
    """

    def __init__(self, *, dtype=np.complex128, noise=None, seed=None, ignore_measurement_results: bool = False):
        self._dtype = dtype
        self._noise = noise
        self._rng = np.random if seed is None else np.random.RandomState(seed)
        self._ignore_measurement_results = ignore_measurement_results

    def _program(self, program: Circuit, param_resolver, qubit_order):
        from . import engine
        resolved = resolve_parameters(program, param_resolver) if param_resolver is not None else program
        return engine.program_for_circuit(resolved, qubit_order)

    def simulate(self, program: Circuit, param_resolver=None, qubit_order=None, initial_state=None) -> DensityMatrixTrialResult:
        if initial_state is not None:
            raise NotImplementedError("only the |0...0> initial state is supported")
        measurements = {}
        if program.has_measurements():
            from . import engine, trajectories
            resolved = resolve_parameters(program, param_resolver) if param_resolver is not None else program
            if self._ignore_measurement_results:
                builder, qubits, meas, _, _ = engine.lower_circuit(resolved, qubit_order, dephase_measurements=True)
                rho = engine.CircuitProgram(builder, qubits=qubits, measurements=meas).density(None)[0]
            else:
                tp, params = trajectories.program_for_circuit(resolved, qubit_order)
                rho, measurements = tp.simulate(self._rng, params)
                qubits = tp.qubits
        else:
            prog, params = self._program(program, param_resolver, qubit_order)
            rho, qubits = prog.density(params)[0], prog.qubits
        if np.dtype(self._dtype) != np.complex128:
            rho = rho.astype(self._dtype)
        result = DensityMatrixTrialResult(rho, qubits, ParamResolver(param_resolver))
        result.measurements = measurements
        return result

    def run(self, program: Circuit, param_resolver=None, repetitions: int = 1) -> TrialResult:
        if not program.has_measurements():
            raise ValueError("Circuit has no measurements to sample.")
        keys = [op.gate.key for op in program.all_operations() if isinstance(op.gate, MeasurementGate)]
        repeated = sorted({k for k in keys if keys.count(k) > 1})
        if repeated:       # cirq: SimulatesSamples.run_sweep -> _verify_unique_measurement_keys
            raise ValueError("Measurement key {} repeated".format(",".join(repeated)))
        if not program.are_all_measurements_terminal():
            from . import trajectories
            resolved = resolve_parameters(program, param_resolver) if param_resolver is not None else program
            tp, params = trajectories.program_for_circuit(resolved, None)
            return TrialResult(tp.run(repetitions, self._rng, params), repetitions)
        prog, params = self._program(program, param_resolver, None)
        probs = prog.probabilities(params, renormalise=True)[0]
        return sample_measurements(probs, prog, repetitions, self._rng)


def sample_measurements(probs: np.ndarray, prog, repetitions: int, rng=np.random) -> TrialResult:
    """Multinomial sampling of terminal measurements from diag(rho) (cirq: evolve once, sample shots)."""
    n = prog.n_qubits
    outcomes = rng.choice(len(probs), size=repetitions, p=probs)
    measurements = {}
    for key, qubit_positions, invert in prog.measurements:
        bits = np.stack([(outcomes >> (n - 1 - pos)) & 1 for pos in qubit_positions], axis=1).astype(np.int8)
        for i, inv in enumerate(invert):
            if inv:
                bits[:, i] ^= 1
        measurements[key] = bits
    return TrialResult(measurements, repetitions)


class _Value:
    validate_probability = staticmethod(validate_probability)


value = _Value()


class _Protocols:
    unitary = staticmethod(unitary)
    mixture = staticmethod(mixture)
    channel = staticmethod(channel)
    resolve_parameters = staticmethod(resolve_parameters)
    is_parameterized = staticmethod(is_parameterized)


protocols = _Protocols()


# --------------------------------------------------------------------------------------------------
# remaining names the reference touches (processor_quantum.py:48-56, error_mitigation.py:743-791 and its type annotations)
# --------------------------------------------------------------------------------------------------


def _with_noise(program: Circuit, noise) -> Circuit:
    """The circuit a ``NoiseModel`` turns ``program`` into: every non-measurement operation through ``noisy_operation``
    (INoiseWrapper is the reference's NoiseModel, i_noise_wrapper.py:73-119).  Measurements stay as they are so that terminal
    ones remain terminal; cirq 0.8.2's default ``noisy_moment`` hands measurement operations to ``noisy_operation`` too (from
    memory - cirq is not installable here), which would append a channel behind the read-out.  The reference never calls the
    two legacy helpers that reach this (``QPU.get_noisy_trial_result`` has no caller), so the difference is unpinned."""
    if noise is None:
        return program
    noisy = Circuit()
    for op in program.all_operations():
        noisy.append(op if isinstance(op.gate, MeasurementGate) else noise.noisy_operation(op))
    return noisy


def sample(program: Circuit, *, noise=None, param_resolver=None, repetitions: int = 1, dtype=np.complex128) -> TrialResult:
    """``cirq.sample`` (reference: ``QPU.get_noisy_trial_result``, processor_quantum.py:55-56): run ``program`` - with the
    noise model applied when one is given - and return the sampled measurement records."""
    return DensityMatrixSimulator(dtype=dtype).run(_with_noise(program, noise), param_resolver=param_resolver, repetitions=repetitions)


class WaveFunctionTrialResult:
    def __init__(self, final_state: np.ndarray, qubits: Sequence[Qid], params: ParamResolver, measurements: Dict[str, np.ndarray]):
        self.final_state = final_state
        self.qubits = list(qubits)
        self.params = params
        self.measurements = measurements

    def state_vector(self) -> np.ndarray:
        return self.final_state


class Simulator:
    """``cirq.Simulator`` (reference: ``QPU.get_trial_results``, processor_quantum.py:48-52).  The engine has one state
    representation, so a noiseless run samples the density-matrix evolution (same distribution); ``simulate`` returns the state
    vector of the pure final state with the global phase fixed by a real, positive largest amplitude (cirq keeps the phase the
    gates accumulate: amplitudes agree up to that one factor).  Circuits with channels belong to DensityMatrixSimulator."""

    def __init__(self, *, dtype=np.complex128, seed=None):
        self._dtype = dtype
        self._dm = DensityMatrixSimulator(dtype=np.complex128, seed=seed)

    def run(self, program: Circuit, param_resolver=None, repetitions: int = 1) -> TrialResult:
        return self._dm.run(program, param_resolver=param_resolver, repetitions=repetitions)

    def simulate(self, program: Circuit, param_resolver=None, qubit_order=None, initial_state=None) -> WaveFunctionTrialResult:
        res = self._dm.simulate(program, param_resolver=param_resolver, qubit_order=qubit_order, initial_state=initial_state)
        rho = res.final_density_matrix
        if abs(np.real(np.trace(rho @ rho)) - np.real(np.trace(rho)) ** 2) > 1e-9:
            raise ValueError("cirq.Simulator: the final state is mixed (the circuit has channels) - use DensityMatrixSimulator")
        k = int(np.argmax(np.real(np.diag(rho))))
        psi = rho[:, k] / np.sqrt(np.real(rho[k, k]))
        return WaveFunctionTrialResult(psi.astype(self._dtype), res.qubits, res.params, getattr(res, "measurements", {}))


class CircuitDiagramInfo:
    def __init__(self, wire_symbols, exponent=1, connected: bool = True):
        self.wire_symbols = tuple(wire_symbols)
        self.exponent = exponent
        self.connected = connected


class CircuitDiagramInfoArgs:
    def __init__(self, known_qubits=None, known_qubit_count=None, use_unicode_characters: bool = True, precision=3, qubit_map=None):
        self.known_qubits, self.known_qubit_count = known_qubits, known_qubit_count
        self.use_unicode_characters, self.precision, self.qubit_map = use_unicode_characters, precision, qubit_map


NOISE_MODEL_LIKE = object          # typing alias in cirq ('cirq.NOISE_MODEL_LIKE' appears in an annotation only)


class _Namespace:
    def __init__(self, **names):
        self.__dict__.update(names)


ops = _Namespace(gate_features=_Namespace(SingleQubitGate=SingleQubitGate, TwoQubitGate=TwoQubitGate, ThreeQubitGate=ThreeQubitGate),
                 pauli_gates=_Namespace(X=X, Y=Y, Z=Z), identity=_Namespace(I=I),
                 SingleQubitGate=SingleQubitGate, TwoQubitGate=TwoQubitGate, ThreeQubitGate=ThreeQubitGate, Gate=Gate,
                 GateOperation=GateOperation, MeasurementGate=MeasurementGate, X=X, Y=Y, Z=Z, I=I, H=H, CNOT=CNOT)
circuits = _Namespace(circuit=Circuit, Circuit=Circuit, Moment=Moment, InsertStrategy=InsertStrategy)
study = _Namespace(resolver=ParamResolver, ParamResolver=ParamResolver, TrialResult=TrialResult)
protocols.CircuitDiagramInfo = CircuitDiagramInfo
protocols.CircuitDiagramInfoArgs = CircuitDiagramInfoArgs
