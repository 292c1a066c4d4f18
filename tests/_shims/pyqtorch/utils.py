from enum import Enum

import torch


class DropoutMode(str, Enum):
    ROTATIONAL = "rotational_dropout"
    ENTANGLING = "entangling_dropout"
    CANONICAL_FWD = "canonical_fwd_dropout"


class SolverType(str, Enum):
    DP5_SE = "dp5_se"
    DP5_ME = "dp5_me"
    KRYLOV_SE = "krylov_se"


class DensityMatrix(torch.Tensor):
    pass


def dm_partial_trace(*a, **k):
    raise RuntimeError("pyqtorch is not installed; tests/_shims provides names only")
