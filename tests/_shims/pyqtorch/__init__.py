"""Name-only stand-in for pyqtorch (absent in this image). No arithmetic."""
from __future__ import annotations


class _Unavailable:
    def __init__(self, *a, **k):
        raise RuntimeError("pyqtorch is not installed; tests/_shims provides names only")


class QuantumCircuit(_Unavailable):
    pass


class DropoutQuantumCircuit(QuantumCircuit):
    pass


Observable = _Unavailable
Scale = Sequence = Add = Merge = HamiltonianEvolution = _Unavailable


def inner_prod(*a, **k):
    raise RuntimeError("pyqtorch is not installed; tests/_shims provides names only")


from . import apply, differentiation, embed, noise, primitives, utils  # noqa: E402,F401
