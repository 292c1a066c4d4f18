"""CPU restatement of the device-side parameter embedding — TEST INFRASTRUCTURE ONLY (never imported by the product).

Interprets an `EmbedProgram` (qadence_b200/embedding.py) with plain torch float64 ops, so that its derivatives come
from torch autograd rather than from the hand-written rules of the CUDA kernel.  It is pinned against the reference's
own host embedding (qadence/blocks/embedding.py:118-167, sympytorch modules) in tests/test_embedding.py.
"""

from __future__ import annotations

import torch
from torch import Tensor

from qadence_b200.embedding import (EO_ABS, EO_ACOS, EO_ADD, EO_ASIN, EO_ATAN, EO_CONST, EO_COS, EO_EXP, EO_FEAT, EO_HEAVISIDE, EO_LOG,
                                    EO_MUL, EO_POW, EO_SIN, EO_TAN, EO_TANH, EO_VAR, EmbedProgram)

_UNARY = {EO_SIN: torch.sin, EO_COS: torch.cos, EO_TAN: torch.tan, EO_TANH: torch.tanh, EO_LOG: torch.log, EO_EXP: torch.exp,
          EO_ACOS: torch.acos, EO_ASIN: torch.asin, EO_ATAN: torch.atan, EO_ABS: torch.abs}


def evaluate(prog: EmbedProgram, var: Tensor, feat: Tensor, batch: int) -> Tensor:
    """[n_slots, batch] float64; differentiable w.r.t. var [n_var] and feat [n_feat, batch] through torch autograd."""
    rows = []
    for s in range(prog.n_slots):
        lo, hi = int(prog.slot_off[s]), int(prog.slot_off[s + 1])
        val: list[Tensor] = []
        for q in prog.nodes[lo:hi]:
            op, a, b, arg = int(q["op"]), int(q["a"]), int(q["b"]), int(q["arg"])
            if op == EO_CONST:
                r = torch.full((batch,), float(prog.consts[arg]), dtype=torch.float64)
            elif op == EO_VAR:
                r = var[arg].to(torch.float64).expand(batch)
            elif op == EO_FEAT:
                r = feat[arg].to(torch.float64)
            elif op == EO_ADD:
                r = val[a] + val[b]
            elif op == EO_MUL:
                r = val[a] * val[b]
            elif op == EO_POW:
                r = torch.pow(val[a], val[b])
            elif op == EO_HEAVISIDE:
                r = torch.heaviside(val[a].detach(), torch.tensor(0.5, dtype=torch.float64))  # as parameters.py:196-205
            else:
                r = _UNARY[op](val[a])
            val.append(r)
        rows.append(val[-1] if val else torch.zeros(batch, dtype=torch.float64))
    return torch.stack(rows) if rows else torch.zeros(0, batch, dtype=torch.float64)
