def count_qubits(*a, **k):
    raise RuntimeError("openfermion is not installed; tests/_shims provides names only")
