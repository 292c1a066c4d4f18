"""BASELINE.json configs at FULL size, checked through size-independent properties (the oracle cannot
finish these sizes): unitarity / round trips, linearity of expectations, adjoint gradients against the
exact parameter-shift rule, sampling consistency, and a slice of cfg3 against the dense-expm oracle."""

from __future__ import annotations

import math

import numpy as np
import pytest
import torch

from oracle import statevec as oracle
from qadence_b200 import engine
from qadence_b200.program import Op, OpKind, PauliSum, PauliTerm, Program, ProgramBuilder, hea_program, total_magnetization

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _inverse(prog: Program) -> Program:
    ops = []
    for op in reversed(prog.ops):
        if op.kind in (OpKind.RX, OpKind.RY, OpKind.RZ, OpKind.PHASE):
            ops.append(Op(op.kind, op.targets, op.controls, param=op.param + "__neg", label=op.label))
        else:
            ops.append(op)  # CNOT is self-inverse
    return Program(prog.n_qubits, ops)


@pytest.fixture(scope="module")
def cfg2():
    n, depth, B = 20, 8, 256
    prog = hea_program(n, depth)
    g = torch.Generator().manual_seed(2)
    vals = {nm: 2 * math.pi * torch.rand(B, generator=g, dtype=torch.float64) for nm in prog.param_names}
    return n, B, prog, vals


def test_cfg2_norm_and_round_trip(cfg2) -> None:
    """configs[1]: hea(20, 8), batch 256, complex128 — U preserves the norm and U†U|0> = |0> to 1e-10."""
    n, B, prog, vals = cfg2
    cp = engine.CompiledProgram(prog)
    psi = engine.run(cp, vals, device=DEV)
    nrm = engine.dot(psi, psi).real
    assert (nrm - 1).abs().max().item() < 1e-12
    del psi
    both = Program(n, prog.ops + _inverse(prog).ops)
    allv = {**vals, **{k + "__neg": -v for k, v in vals.items()}}
    back = engine.run(engine.CompiledProgram(both), allv, device=DEV)
    assert (back[:, 0] - 1).abs().max().item() < 1e-10
    back[:, 0] = 0
    assert back.abs().max().item() < 1e-10


def test_cfg2_adjoint_gradients_equal_parameter_shift(cfg2) -> None:
    """dE/dθ for RX / RY is exactly (E(θ+π/2) − E(θ−π/2)) / 2; checked at full size for first / middle / last gates."""
    n, B, prog, vals = cfg2
    cp = engine.CompiledProgram(prog)
    cob = engine.CompiledObservable(total_magnetization(n))
    leaf = {k: v.clone().requires_grad_(True) for k, v in vals.items()}
    e = engine.expectation(cp, [cob], leaf, device=DEV)
    e.sum().backward()
    names = prog.param_names
    for nm in (names[0], names[len(names) // 2 + 7], names[-1]):
        plus = {**vals, nm: vals[nm] + math.pi / 2}
        minus = {**vals, nm: vals[nm] - math.pi / 2}
        with torch.no_grad():
            ep = engine.expectation(cp, [cob], plus, device=DEV)[:, 0]
            em = engine.expectation(cp, [cob], minus, device=DEV)[:, 0]
        psr = ((ep - em) / 2).cpu()
        assert (leaf[nm].grad - psr).abs().max().item() < 1e-10, nm


def test_cfg2_expectation_linearity_and_sampling(cfg2) -> None:
    n, B, prog, vals = cfg2
    cp = engine.CompiledProgram(prog)
    o1 = total_magnetization(n)
    o2 = PauliSum(n, [PauliTerm(0.7, (), ((0, "X"), (n - 1, "Y"))), PauliTerm(-0.4, (), ((3, "Z"), (11, "Z")))])
    o12 = PauliSum(n, o1.terms + o2.terms)
    e = engine.expectation(cp, [engine.CompiledObservable(o) for o in (o1, o2, o12)], vals, device=DEV)
    assert (e[:, 0] + e[:, 1] - e[:, 2]).abs().max().item() < 1e-10
    assert e[:, 0].abs().max().item() <= n + 1e-9
    # sampled <Z_tot> of one batch element agrees with the exact value within 6 standard errors
    sub = {k: v[:2] for k, v in vals.items()}
    psi = engine.run(cp, sub, device=DEV)
    shots = 20000
    u = torch.rand(2, shots, dtype=torch.float64, generator=torch.Generator().manual_seed(5))
    idx = engine.sample_indices(psi, u).cpu().numpy()
    ztot = np.array([[n - 2 * bin(int(i)).count("1") for i in row] for row in idx])
    exact = e[:2, 0].cpu().numpy()
    assert np.all(np.abs(ztot.mean(1) - exact) < 6 * ztot.std(1) / math.sqrt(shots) + 1e-9)


def test_cfg4_training_step_properties() -> None:
    """configs[3] per-GPU share: feature map + hea(16, 12), 512 samples: loss gradient of a shared θ equals the sum of
    per-sample parameter-shift gradients."""
    n, B = 16, 512
    b = ProgramBuilder(n)
    for q in range(n):
        b.RX(q, "x")
    prog = Program(n, b.build().ops + hea_program(n, 12).ops)
    cp = engine.CompiledProgram(prog)
    cob = engine.CompiledObservable(total_magnetization(n))
    g = torch.Generator().manual_seed(4)
    theta = {nm: (2 * math.pi * torch.rand(1, generator=g, dtype=torch.float64)).requires_grad_(True) for nm in prog.param_names if nm != "x"}
    x = torch.rand(B, generator=g, dtype=torch.float64) * 1.98 - 0.99
    e = engine.expectation(cp, [cob], {**theta, "x": x}, device=DEV)
    e.sum().backward()
    nm = "theta_300"
    with torch.no_grad():
        base = {k: v.detach() for k, v in theta.items()}
        ep = engine.expectation(cp, [cob], {**base, nm: base[nm] + math.pi / 2, "x": x}, device=DEV)
        em = engine.expectation(cp, [cob], {**base, nm: base[nm] - math.pi / 2, "x": x}, device=DEV)
    assert abs(theta[nm].grad.item() - ((ep - em) / 2).sum().item()) < 1e-9


def test_cfg3_full_batch_norm_and_slice_parity() -> None:
    """configs[2]: 12-atom Rydberg register (3x4 lattice, 8 um), AnalogRX / AnalogInteraction / AnalogRY as HamEvo with the
    generators `analog/hamiltonian_terms.py:16-72` produces, batch 1024: norm for the whole batch, one element against the
    dense 4096 x 4096 matrix_exp oracle."""
    n, B = 12, 1024
    coords = [(8.0 * (i // 4), 8.0 * (i % 4)) for i in range(n)]
    c6 = 865723.02
    tint = []
    for i in range(n):
        for j in range(i + 1, n):
            r = math.hypot(coords[i][0] - coords[j][0], coords[i][1] - coords[j][1])
            c = c6 / r**6 / 4.0
            tint += [PauliTerm(c, (), ()), PauliTerm(-c, (), ((i, "Z"),)), PauliTerm(-c, (), ((j, "Z"),)), PauliTerm(c, (), ((i, "Z"), (j, "Z")))]
    h_int = PauliSum(n, tint).simplified()
    om = math.pi
    h_rx = PauliSum(n, h_int.terms + [PauliTerm(om / 2, (), ((q, "X"),)) for q in range(n)]).simplified()
    h_ry = PauliSum(n, h_int.terms + [PauliTerm(om / 2, (), ((q, "Y"),)) for q in range(n)]).simplified()
    prog = ProgramBuilder(n).hamevo(h_rx, "ta").hamevo(h_int, "tt").hamevo(h_ry, "tb").build()
    g = torch.Generator().manual_seed(0)
    vals = {"ta": torch.rand(B, generator=g, dtype=torch.float64), "tt": torch.rand(B, generator=g, dtype=torch.float64),
            "tb": torch.rand(B, generator=g, dtype=torch.float64)}
    psi = engine.run(engine.CompiledProgram(prog), vals, device=DEV)
    assert (engine.dot(psi, psi).real - 1).abs().max().item() < 1e-11
    want = oracle.run(prog, {k: v[:1] for k, v in vals.items()})
    assert (psi[:1].cpu() - want).abs().max().item() < 1e-10
