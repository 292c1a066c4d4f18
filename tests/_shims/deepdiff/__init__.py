class DeepDiff(dict):
    def __init__(self, a, b, **k):
        super().__init__({} if a == b else {"values_changed": 1})
