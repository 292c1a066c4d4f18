"""`BackendName` / `Engine` with the b200sv entries (picked up by qadence/types.py:244-252).

Imported while `qadence.types` is still initialising, so it must not import qadence."""

from __future__ import annotations

from enum import Enum


class StrEnum(str, Enum):
    def __str__(self) -> str:
        return str(self.value)

    @classmethod
    def list(cls) -> list[str]:
        return [c.value for c in cls]


class BackendName(StrEnum):
    PYQTORCH = "pyqtorch"
    PULSER = "pulser"
    HORQRUX = "horqrux"
    B200SV = "b200sv"
    """B200-native state-vector backend (hand-written sm_100a CUDA)."""


class Engine(StrEnum):
    TORCH = "torch"
    JAX = "jax"
    B200SV = "b200sv"
    """torch autograd front + the backend's own fused adjoint pass."""
