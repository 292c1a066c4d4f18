"""`BackendName` / `Engine` with the b200sv entries (picked up by qadence/types.py:244-252).

Imported while `qadence.types` is still initialising, so it must not import qadence."""

from __future__ import annotations

import os
from enum import Enum

# Replace mode (`B200SV_REPLACE_PYQTORCH=1`, see qadence_extensions/__init__.py): b200sv answers to the NAME "pyqtorch", so
# the enums stay exactly qadence's own (code that iterates over BackendName.list() sees nothing new).
REPLACE_PYQTORCH = os.environ.get("B200SV_REPLACE_PYQTORCH", "0") not in ("", "0")


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
    if not REPLACE_PYQTORCH:
        B200SV = "b200sv"
        """B200-native state-vector backend (hand-written sm_100a CUDA)."""


class Engine(StrEnum):
    TORCH = "torch"
    JAX = "jax"
    if not REPLACE_PYQTORCH:
        B200SV = "b200sv"
        """torch autograd front + the backend's own fused adjoint pass."""
