"""``IParameter``: the ordered name -> value table of an ansatz (API of
``src/data_containers/helper_interfaces/i_parameter.py``).

One table object is shared BY REFERENCE between an ansatz and its noise wrapper (the reference's
``i_noise_wrapper.py:100`` hands the same object to both and callers rely on that aliasing), so every mutator works in
place on the wrapped mapping and never rebinds it.
"""
from __future__ import annotations

from typing import Dict, Iterator, Sequence

from ... import cirq_compat as cirq


class IParameter:
    def __init__(self, parameters: Dict[str, float]):
        self._p = parameters

    # ---- mapping protocol --------------------------------------------------------------------------------
    def __getitem__(self, name: str) -> float:
        return self._p[name]

    def __setitem__(self, name: str, value: float) -> None:
        self._p[name] = value

    def __iter__(self) -> Iterator[str]:
        yield from self._p

    def __len__(self) -> int:
        return len(self._p)

    def __str__(self) -> str:
        return str(self._p)

    def __add__(self, other: "IParameter") -> "IParameter":
        """New table holding this one's entries followed by ``other``'s (``other`` wins on equal names)."""
        combined = dict(self._p)
        combined.update(other.get_dict())
        return IParameter(combined)

    # ---- reference API -------------------------------------------------------------------------------------
    def get_dict(self) -> Dict[str, float]:
        """The wrapped mapping itself (not a copy)."""
        return self._p

    dict = property(get_dict)

    def get_resolved(self) -> cirq.ParamResolver:
        """Resolver over the current values (reference :8-9)."""
        return cirq.ParamResolver(self._p)

    def update(self, v: Sequence[float]) -> None:
        """Positional assignment of optimiser output, in the table's own order (reference :24-28)."""
        names = list(self._p)
        if len(names) != len(v):
            raise IndexError("Dimensions of optimized trial result does not correspond to IParameter")
        for name, value in zip(names, v):
            self._p[name] = value
