from enum import Enum


class WhiteNoise(Enum):
    UNIFORM = "uniform"
    GAUSSIAN = "gaussian"
    POISSON = "poisson"
