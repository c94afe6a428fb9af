"""Reader / writer for the reference's on-disk result format (``jsonpickle==1.4.1`` JSON, ``requirements.txt``).

The reference persists its experiment results with ``jsonpickle.encode(value=obj, indent=4)``
(``src/calculate_minimisation.py:38-71``) and reads them back for plotting (``src/plot_minimisation.py:9-17``,
``src/visual_testing.py:866-871``).  jsonpickle is not installed here, so this module restates the slice of
its wire format those files use — nothing more:

* ``{"py/object": "module.Class", <attributes in __dict__ order>}``;
* ``{"py/reduce": [f, args, state]}`` for numpy scalars (``numpy.core.multiarray.scalar``), arrays
  (``numpy.core.multiarray._reconstruct``) and dtypes (``numpy.dtype``), raw bytes as ``{"py/b64": ...}``;
* ``{"py/tuple": [...]}``, ``{"py/type": ...}``, ``{"py/function": ...}``;
* ``{"py/id": k}``: back-reference to the k-th (zero-based) object in first-visit order, where lists, class
  instances and reduce-objects are counted and tuples, dicts, bytes and primitives are not.

``encode`` writes classes of this package under the reference's module names (``src.…``) and ``decode`` maps
them back, so files written here are readable by the reference's scripts and vice versa.  Callables named by
``py/function`` / ``py/type`` are resolved through a fixed whitelist: decoding never imports or calls
anything else (unlike jsonpickle itself).
"""
from __future__ import annotations

import base64
import importlib
import json
from typing import Any, Dict, List

import numpy as np

OBJECT, REDUCE, TUPLE, TYPE, FUNCTION, ID, B64 = "py/object", "py/reduce", "py/tuple", "py/type", "py/function", "py/id", "py/b64"

_PKG = __name__.rsplit(".", 1)[0]
_REF_ROOT = "src"

_SCALAR_FN = "numpy.core.multiarray.scalar"
_RECONSTRUCT_FN = "numpy.core.multiarray._reconstruct"
_DTYPE_TYPE = "numpy.dtype"
_NDARRAY_TYPE = "numpy.ndarray"


class JsonPickleFormatError(ValueError):
    pass


# ------------------------------------------------------------------------------------------------- encoding

def _module_to_reference(module: str) -> str:
    if module == _PKG or module.startswith(_PKG + "."):
        return _REF_ROOT + module[len(_PKG):]
    return module


def _module_from_reference(module: str) -> str:
    if module == _REF_ROOT or module.startswith(_REF_ROOT + "."):
        return _PKG + module[len(_REF_ROOT):]
    return module


def _dtype_state(dt: np.dtype) -> list:
    # dtype.__reduce__() state of a plain (non-structured) dtype, as numpy 1.19 wrote it
    order = "|" if dt.itemsize == 1 or dt.byteorder == "|" else "<"
    return [3, order, None, None, None, -1, -1, 0]


class _Encoder:
    def __init__(self):
        self._ids: Dict[int, int] = {}
        self._keep: List[Any] = []          # keeps id()-tracked objects alive while encoding
        self._scalar_dtypes: Dict[str, int] = {}

    def _new_id(self, obj=None) -> int:
        k = len(self._keep)
        self._keep.append(obj)
        if obj is not None:
            self._ids[id(obj)] = k
        return k

    def _dtype(self, dt: np.dtype, shared: bool):
        key = dt.str
        if shared and key in self._scalar_dtypes:
            return {ID: self._scalar_dtypes[key]}
        k = self._new_id()
        if shared:
            self._scalar_dtypes[key] = k
        return {REDUCE: [{TYPE: _DTYPE_TYPE}, {TUPLE: [dt.str.lstrip("<|=>"), False, True]}, {TUPLE: _dtype_state(dt)}]}

    def flatten(self, obj):
        if obj is None or (isinstance(obj, (bool, int, float, str)) and not isinstance(obj, np.generic)):
            return obj
        if isinstance(obj, bytes):
            return {B64: base64.b64encode(obj).decode("ascii")}
        if id(obj) in self._ids:
            return {ID: self._ids[id(obj)]}
        if isinstance(obj, np.generic):
            self._new_id(obj)
            dt = self._dtype(obj.dtype, shared=True)
            return {REDUCE: [{FUNCTION: _SCALAR_FN}, {TUPLE: [dt, self.flatten(obj.tobytes())]}]}
        if isinstance(obj, np.ndarray):
            if obj.dtype.hasobject or obj.dtype.fields is not None:
                raise JsonPickleFormatError("object / structured arrays are not part of the stored format")
            self._new_id(obj)
            args = {TUPLE: [{TYPE: _NDARRAY_TYPE}, {TUPLE: [0]}, {B64: "Yg=="}]}
            fortran = bool(obj.ndim > 1 and obj.flags.f_contiguous and not obj.flags.c_contiguous)
            raw = obj.tobytes(order="F" if fortran else "C")
            state = {TUPLE: [1, {TUPLE: [int(s) for s in obj.shape]}, self._dtype(obj.dtype, shared=False), fortran, self.flatten(raw)]}
            return {REDUCE: [{FUNCTION: _RECONSTRUCT_FN}, args, state]}
        if isinstance(obj, list):
            self._new_id(obj)
            return [self.flatten(v) for v in obj]
        if isinstance(obj, tuple):
            return {TUPLE: [self.flatten(v) for v in obj]}
        if isinstance(obj, dict):
            out = {}
            for k, v in obj.items():
                if not isinstance(k, str):
                    raise JsonPickleFormatError("only string dictionary keys are part of the stored format")
                out[k] = self.flatten(v)
            return out
        if isinstance(obj, type):
            return {TYPE: f"{_module_to_reference(obj.__module__)}.{obj.__qualname__}"}
        if hasattr(obj, "__dict__"):
            self._new_id(obj)
            cls = type(obj)
            out = {OBJECT: f"{_module_to_reference(cls.__module__)}.{cls.__qualname__}"}
            for k, v in vars(obj).items():
                out[k] = self.flatten(v)
            return out
        raise JsonPickleFormatError(f"cannot encode {type(obj).__name__}")


def encode(value: Any, indent: int | None = None) -> str:
    """``jsonpickle.encode(value=..., indent=...)`` for the container types the reference stores."""
    return json.dumps(_Encoder().flatten(value), indent=indent)


# ------------------------------------------------------------------------------------------------- decoding

class _Pending:
    """Slot of a reduce-object whose value is known only after its arguments are decoded."""
    __slots__ = ("value",)

    def __init__(self):
        self.value = None


def _np_scalar(dtype: np.dtype, raw: bytes):
    return np.frombuffer(raw, dtype=dtype, count=1)[0]


_FUNCTIONS = {
    "numpy.core.multiarray.scalar": _np_scalar,
    "numpy._core.multiarray.scalar": _np_scalar,
    "numpy.core.multiarray._reconstruct": "reconstruct",
    "numpy._core.multiarray._reconstruct": "reconstruct",
}
_TYPES = {"numpy.dtype": np.dtype, "numpy.ndarray": np.ndarray}


class _Decoder:
    def __init__(self):
        self._objs: List[Any] = []

    def _register(self, obj) -> int:
        self._objs.append(obj)
        return len(self._objs) - 1

    def restore(self, node):
        if isinstance(node, list):
            out: list = []
            self._register(out)
            out.extend(self.restore(v) for v in node)
            return out
        if not isinstance(node, dict):
            return node
        if ID in node:
            k = node[ID]
            if not isinstance(k, int) or not 0 <= k < len(self._objs):
                raise JsonPickleFormatError(f"py/id {k!r} refers to nothing decoded so far")
            obj = self._objs[k]
            return obj.value if isinstance(obj, _Pending) else obj
        if B64 in node:
            return base64.b64decode(node[B64])
        if TUPLE in node:
            return tuple(self.restore(v) for v in node[TUPLE])
        if TYPE in node:
            return self._named(node[TYPE], _TYPES)
        if FUNCTION in node:
            return self._named(node[FUNCTION], _FUNCTIONS)
        if REDUCE in node:
            return self._reduce(node[REDUCE])
        if OBJECT in node:
            return self._instance(node)
        return {k: self.restore(v) for k, v in node.items()}

    @staticmethod
    def _named(name: str, table: dict):
        if name not in table:
            raise JsonPickleFormatError(f"{name!r} is not a callable / type the stored format uses")
        return table[name]

    def _reduce(self, parts: list):
        slot = _Pending()
        k = self._register(slot)
        vals = [self.restore(p) for p in parts] + [None] * (3 - len(parts))
        f, args, state = vals[0], vals[1] or (), vals[2]
        if f is np.dtype:
            value = np.dtype(args[0])
            if state is not None and len(state) > 1 and state[1] in (">",):
                value = value.newbyteorder(">")
        elif f == "reconstruct":
            if state is None or len(state) < 5:
                raise JsonPickleFormatError("ndarray without state")
            _ver, shape, dtype, fortran, raw = state[:5]
            value = np.frombuffer(raw, dtype=dtype).reshape(shape, order="F" if fortran else "C").copy()
        elif callable(f):
            value = f(*args)
        else:
            raise JsonPickleFormatError("unsupported py/reduce")
        slot.value = value
        self._objs[k] = value
        return value

    def _instance(self, node: dict):
        module, _, name = node[OBJECT].rpartition(".")
        module = _module_from_reference(module)
        if not (module == _PKG or module.startswith(_PKG + ".")):
            raise JsonPickleFormatError(f"{node[OBJECT]!r}: only this package's container classes are decoded")
        cls = getattr(importlib.import_module(module), name)
        obj = cls.__new__(cls)
        self._register(obj)
        for k, v in node.items():
            if k != OBJECT:
                obj.__dict__[k] = self.restore(v)
        return obj


def decode(string: str) -> Any:
    """``jsonpickle.decode(string)`` for files written by the reference (or by :func:`encode`)."""
    return _Decoder().restore(json.loads(string))
