from qadence_b200.qadence_backend.config import Configuration  # noqa: F401
