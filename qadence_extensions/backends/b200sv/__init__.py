from qadence_b200.qadence_backend import Backend, Configuration, supported_gates  # noqa: F401
