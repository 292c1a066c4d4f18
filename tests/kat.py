"""Helpers shared by the oracle (CPU) and CUDA (GPU) known-answer tests."""

from __future__ import annotations

import json
from pathlib import Path

import torch

from qadence_b200.program import Program, ProgramBuilder

GOLDEN = Path(__file__).resolve().parent / "golden"


def load_kats() -> dict:
    return json.loads((GOLDEN / "kat_reference.json").read_text())


def kat_program(n: int, ops: list) -> Program:
    b = ProgramBuilder(n)
    for op in ops:
        name, args = op[0], op[1:]
        getattr(b, name)(*args)
    return b.build()


def kat_state(pairs: list) -> torch.Tensor:
    return torch.tensor([[complex(re, im) for re, im in pairs]], dtype=torch.complex128)
