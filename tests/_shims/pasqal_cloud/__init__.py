class SDK:
    pass


class Workload:
    pass
