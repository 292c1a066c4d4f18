"""Functional stand-in for sympytorch.SymPyModule (absent in this image).

Evaluates ONE sympy expression on torch tensors via ``sympy.lambdify``; differentiable.
Host-side parameter-embedding glue only (qadence/parameters.py:208-211 is the caller).
"""
from __future__ import annotations

import sympy
import torch


class SymPyModule(torch.nn.Module):
    def __init__(self, expressions, extra_funcs=None):
        super().__init__()
        (self._expr,) = expressions
        self._syms = sorted(self._expr.free_symbols, key=lambda s: s.name)
        table = {
            "Heaviside": lambda x, *a: torch.heaviside(
                torch.as_tensor(x).detach(), torch.tensor(0.5, dtype=torch.as_tensor(x).dtype)
            ),
            "I": 1j,
        }
        self._fn = sympy.lambdify(self._syms, self._expr, modules=[table, torch])

    def forward(self, **kw):
        out = self._fn(*[torch.as_tensor(kw[s.name]) for s in self._syms])
        out = out if torch.is_tensor(out) else torch.as_tensor(out)
        return out.unsqueeze(-1)
