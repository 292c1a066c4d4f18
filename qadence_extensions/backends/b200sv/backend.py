"""`import_backend("b200sv")` lands here (qadence/extensions.py:47-56)."""
from qadence_b200.qadence_backend.backend import Backend  # noqa: F401
