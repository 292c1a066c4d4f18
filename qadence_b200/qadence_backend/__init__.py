"""qadence-facing mirror of the reference backend interface (importable only where qadence is)."""

from .backend import Backend
from .config import Configuration
from .convert_ops import supported_gate_names

supported_gates = supported_gate_names()

__all__ = ["Backend", "Configuration", "supported_gates"]
