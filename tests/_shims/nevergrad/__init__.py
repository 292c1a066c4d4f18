class _NS:
    pass


optimizers = _NS()
optimizers.Optimizer = object
p = _NS()
p.Array = object
