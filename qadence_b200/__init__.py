"""b200sv — a B200-native (sm_100a) state-vector backend for qadence.

Layout (only what the hot path needs):
  program.py   flat op-program IR (what crosses the drop-in boundary)
  plan.py      fusion planner: gates → sweeps → register rounds
  cabi.py      ctypes binding of the C-ABI library (include/b200sv.h)
  engine.py    run / expectation (+ adjoint autograd) / sample on the GPU
  hamevo.py    HamEvo: on-the-fly diagonal evolution and matrix-free Krylov
  dist.py      batch sharding and global-qubit sharding over NCCL
  csrc/        hand-written CUDA kernels + the C-ABI
  qadence_backend/   the qadence `Backend` / `DifferentiableBackend` / `Configuration` mirror
"""

from .program import (  # noqa: F401
    CONST_MATS,
    Op,
    OpKind,
    PauliSum,
    PauliTerm,
    Program,
    ProgramBuilder,
    hea_program,
    total_magnetization,
)

__version__ = "0.1.0"
BACKEND_NAME = "b200sv"
