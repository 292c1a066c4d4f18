"""BASELINE.json configs[1] and configs[3] at their OWN width against the oracle (file name sorts last: ~1 minute of CPU
work).  The CUDA path runs the full production plans (n = 20 / 16, lean TMA sweeps); a few batch elements drawn from
bench.py's own seed are compared — expectation AND every gradient — with `oracle.adjoint_expectation_and_grads`
(einsum per gate + adjoint loop, the restated pyqtorch path) at 1e-10."""

from __future__ import annotations

import math

import pytest
import torch

from oracle import statevec as oracle
from qadence_b200 import engine
from qadence_b200.program import Program, ProgramBuilder, hea_program, total_magnetization

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_cfg2_expectation_and_all_gradients_match_the_oracle() -> None:
    """hea(20, 8): 3 of the 256 parameter sets bench.py draws (same generator, seed 1234, rank 0), all 480 gradients."""
    n, depth, B = 20, 8, 256
    prog = hea_program(n, depth)
    cp = engine.CompiledProgram(prog)
    obs = total_magnetization(n)
    g = torch.Generator().manual_seed(1234)
    table = 2 * torch.pi * torch.rand(len(cp.names), B, generator=g, dtype=torch.float64)  # bench.py: host_params
    pick = [0, 101, 255]
    vals = {nm: table[i, pick].clone().requires_grad_(True) for i, nm in enumerate(cp.names)}
    out = engine.expectation(cp, [engine.CompiledObservable(obs)], vals, device=DEV)
    out.sum().backward()
    want_e, want_g = oracle.adjoint_expectation_and_grads(prog, obs, {k: v.detach() for k, v in vals.items()})
    assert (out[:, 0].detach().cpu() - want_e).abs().max().item() <= 1e-10
    worst = max((vals[k].grad.cpu() - want_g[k]).abs().max().item() for k in vals)
    assert worst <= 1e-10, worst


def test_cfg2_full_batch_equals_its_slices() -> None:
    """The same three elements evaluated inside the full batch of 256 (the bench's exact launch geometry) give the same
    numbers as evaluated alone: batch elements are independent state vectors."""
    n, depth, B = 20, 8, 256
    prog = hea_program(n, depth)
    cp = engine.CompiledProgram(prog)
    cob = engine.CompiledObservable(total_magnetization(n))
    g = torch.Generator().manual_seed(1234)
    table = 2 * torch.pi * torch.rand(len(cp.names), B, generator=g, dtype=torch.float64)
    full = {nm: table[i].clone().requires_grad_(True) for i, nm in enumerate(cp.names)}
    e = engine.expectation(cp, [cob], full, device=DEV)
    e.sum().backward()
    pick = [0, 101, 255]
    sub = {nm: table[i, pick].clone().requires_grad_(True) for i, nm in enumerate(cp.names)}
    es = engine.expectation(cp, [cob], sub, device=DEV)
    es.sum().backward()
    assert (e[pick, 0] - es[:, 0]).abs().max().item() <= 1e-12
    assert max((full[k].grad[pick] - sub[k].grad).abs().max().item() for k in full) <= 1e-12


def test_cfg4_expectation_and_all_gradients_match_the_oracle() -> None:
    """feature map + hea(16, 12) (772 gates, 576 shared θ + x): 2 samples; dE/dθ (summed over the samples, θ is shared)
    and dE/dx per sample."""
    n = 16
    b = ProgramBuilder(n)
    for q in range(n):
        b.RX(q, "x")
    prog = Program(n, b.build().ops + hea_program(n, 12).ops)
    cp = engine.CompiledProgram(prog)
    obs = total_magnetization(n)
    g = torch.Generator().manual_seed(4)
    theta = {nm: (2 * math.pi * torch.rand(1, generator=g, dtype=torch.float64)).requires_grad_(True) for nm in prog.param_names if nm != "x"}
    x = (torch.rand(2, generator=g, dtype=torch.float64) * 1.98 - 0.99).requires_grad_(True)
    out = engine.expectation(cp, [engine.CompiledObservable(obs)], {**theta, "x": x}, device=DEV)
    out.sum().backward()
    ovals = {**{k: v.detach().expand(2) for k, v in theta.items()}, "x": x.detach()}
    want_e, want_g = oracle.adjoint_expectation_and_grads(prog, obs, ovals)
    assert (out[:, 0].detach().cpu() - want_e).abs().max().item() <= 1e-10
    assert (x.grad - want_g["x"]).abs().max().item() <= 1e-10
    worst = max(abs(theta[k].grad.item() - want_g[k].sum().item()) for k in theta)
    assert worst <= 1e-10, worst
