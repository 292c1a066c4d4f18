class ConcretizedCallable:
    def __init__(self, *a, **k):
        raise RuntimeError("pyqtorch is not installed; tests/_shims provides names only")
