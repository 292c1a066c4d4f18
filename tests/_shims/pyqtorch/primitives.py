class Primitive:
    pass


class Parametric(Primitive):
    pass
