class ParserPEG:
    def __init__(self, *a, **k):
        pass
