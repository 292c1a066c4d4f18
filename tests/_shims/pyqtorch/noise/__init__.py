from enum import Enum


class DigitalNoiseType(str, Enum):
    BITFLIP = "BitFlip"
    PHASEFLIP = "PhaseFlip"
    DEPOLARIZING = "Depolarizing"
    PAULI_CHANNEL = "PauliChannel"
    AMPLITUDE_DAMPING = "AmplitudeDamping"
    PHASE_DAMPING = "PhaseDamping"
    GENERALIZED_AMPLITUDE_DAMPING = "GeneralizedAmplitudeDamping"


class DigitalNoiseProtocol:
    pass


class ReadoutNoise:
    pass


class CorrelatedReadoutNoise:
    pass


from . import readout  # noqa: E402,F401
