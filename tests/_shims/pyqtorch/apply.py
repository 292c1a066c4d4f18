def apply_operator(*a, **k):
    raise RuntimeError("pyqtorch is not installed; tests/_shims provides names only")
