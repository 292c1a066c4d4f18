class QubitOperator:
    terms: dict = {}
