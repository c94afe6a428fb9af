"""Named-parameter container (mirror of ``src/data_containers/helper_interfaces/i_parameter.py``)."""
from __future__ import annotations

from typing import Dict, Sequence

from ... import cirq_compat as cirq


class IParameter:
    """Ordered ``name -> value`` mapping shared *by reference* between an ansatz and its noise wrapper
    (reference: i_noise_wrapper.py:100 hands the same object to both, callers rely on that aliasing)."""

    def __init__(self, parameters: Dict[str, float]):
        self._p = parameters

    def get_dict(self) -> Dict[str, float]:
        return self._p

    dict = property(get_dict)

    def get_resolved(self) -> cirq.ParamResolver:
        """Reference: i_parameter.py:8-9 (cirq.ParamResolver over the dict)."""
        return cirq.ParamResolver(self._p)

    def update(self, v: Sequence[float]):
        if len(v) != len(self._p):
            raise IndexError("Dimensions of optimized trial result does not correspond to IParameter")
        for key, value in zip(list(self._p.keys()), v):
            self._p[key] = value

    def __iter__(self):
        return iter(self._p)

    def __len__(self):
        return len(self._p)

    def __str__(self):
        return str(self._p)

    def __getitem__(self, item: str):
        return self._p[item]

    def __setitem__(self, key: str, value: float):
        self._p[key] = value

    def __add__(self, other: "IParameter") -> "IParameter":
        merged = dict(self._p)
        merged.update(other.dict)
        return IParameter(merged)
