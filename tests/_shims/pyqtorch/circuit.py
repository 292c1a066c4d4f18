"""Name-only stand-in for pyqtorch.circuit."""
from . import DropoutQuantumCircuit, QuantumCircuit  # noqa: F401
