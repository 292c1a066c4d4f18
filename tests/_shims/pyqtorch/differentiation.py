class AdjointExpectation:
    @staticmethod
    def apply(*a, **k):
        raise RuntimeError("pyqtorch is not installed; tests/_shims provides names only")
