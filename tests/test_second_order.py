"""Second- and third-order derivatives of expectation values (DQC-style losses).

The reference gets them from torch autograd through pyqtorch's dense ops (`ml_tools/models.py:27-35`, create_graph=True);
its higher-order adjoint tests are skipped (tests/backends/test_adjoint.py:127-248).  Here the adjoint gradient is itself a
differentiable Function (exact parameter shifts of the gates the cotangent touches).  Oracle: torch autograd through the
CPU restatement, to any order."""

from __future__ import annotations

import numpy as np
import pytest
import sympy
import torch

from oracle import statevec as oracle
from qadence_b200.program import OpKind, PauliSum, PauliTerm, ProgramBuilder, hea_program, total_magnetization

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _fm_hea(n: int):
    b = ProgramBuilder(n)
    for q in range(n):
        b.RX(q, f"fm{q}")
    prog = b.build()
    prog.ops += hea_program(n, 1).ops
    theta = [nm for nm in prog.param_names if not nm.startswith("fm")]
    return prog, theta


def _values(n, theta, x, th):
    vals = {f"fm{q}": (q + 1) * torch.acos(x) for q in range(n)}
    vals.update({nm: th[i].reshape(1) for i, nm in enumerate(theta)})
    return vals


def _derivs(expect, x, th):
    """(dE/dx, d2E/dx2, d/dtheta of sum(dE/dx)) of the summed expectation."""
    e = expect()
    (dx,) = torch.autograd.grad(e.sum(), x, create_graph=True)
    dxx, dxt = torch.autograd.grad(dx.sum(), (x, th), create_graph=False)
    return e.detach().cpu(), dx.detach().cpu(), dxx.cpu(), dxt.cpu()


def test_dqc_derivatives_host_dict() -> None:
    from qadence_b200 import engine

    n, B = 3, 4
    prog, theta = _fm_hea(n)
    obs = PauliSum(n, [PauliTerm(1.0, (), ((0, "Z"),)), PauliTerm(0.7, (), ((1, "X"), (2, "Y")))])
    rng = np.random.default_rng(0)
    x0, t0 = rng.uniform(-0.8, 0.8, size=B), rng.uniform(0, 3, size=len(theta))
    xh, th = torch.tensor(x0, requires_grad=True), torch.tensor(t0, requires_grad=True)
    want = _derivs(lambda: oracle.expectation(prog, [obs], _values(n, theta, xh, th)), xh, th)
    xd, td = torch.tensor(x0, device=DEV, requires_grad=True), torch.tensor(t0, device=DEV, requires_grad=True)
    cp, cob = engine.CompiledProgram(prog), engine.CompiledObservable(obs)
    got = _derivs(lambda: engine.expectation(cp, [cob], _values(n, theta, xd, td), device=DEV), xd, td)
    for g, w, name in zip(got, want, ("E", "dE/dx", "d2E/dx2", "d2E/dxdtheta")):
        assert (g - w).abs().max() < 1e-9, (name, g, w)


def test_dqc_loss_through_device_embedding_and_third_order() -> None:
    from qadence_b200 import embedding as E
    from qadence_b200 import engine

    n, B = 2, 3
    prog, theta = _fm_hea(n)
    obs = total_magnetization(n)
    xs = sympy.Symbol("x")
    eprog = E.compile_expressions([(f"fm{q}", (q + 1) * sympy.acos(xs)) for q in range(n)] + [(nm, sympy.Symbol(nm)) for nm in theta],
                                  lambda nm: nm == "x")
    rng = np.random.default_rng(1)
    x0, t0 = rng.uniform(-0.7, 0.7, size=B), rng.uniform(0, 3, size=len(theta))
    cp, cob = engine.CompiledProgram(prog), engine.CompiledObservable(obs)

    def loss_fn(expect, x, th):
        e = expect()[:, 0]
        (dx,) = torch.autograd.grad(e.sum(), x, create_graph=True)
        (dxx,) = torch.autograd.grad(dx.sum(), x, create_graph=True)
        return ((dxx + 0.3 * dx - e) ** 2).mean()  # an ODE residual containing first and second derivatives

    xh, th = torch.tensor(x0, requires_grad=True), torch.tensor(t0, requires_grad=True)
    lw = loss_fn(lambda: oracle.expectation(prog, [obs], _values(n, theta, xh, th)), xh, th)
    gw = torch.autograd.grad(lw, th)[0]

    xd, td = torch.tensor(x0, device=DEV, requires_grad=True), torch.tensor(t0, device=DEV, requires_grad=True)
    order = [theta.index(nm) for nm in eprog.var_names]

    def expect_dev():
        table = E.evaluate(eprog, td[order], xd.reshape(1, B), B)
        return engine.expectation(cp, [cob], E.ParamTable(eprog.names, table, {}, torch.device("cpu")), device=DEV)

    lg = loss_fn(expect_dev, xd, td)
    gg = torch.autograd.grad(lg, td)[0]  # third-order mixed derivatives of E
    assert abs(lg.item() - lw.item()) < 1e-9
    assert (gg.cpu() - gw).abs().max() < 1e-8


def test_controlled_rotation_uses_the_four_term_rule() -> None:
    from qadence_b200 import engine

    n, B = 3, 2
    prog = (ProgramBuilder(n).H(0).RY(1, "a").rot(OpKind.RX, 2, "c", controls=(0,)).rot(OpKind.RZ, 1, "d", controls=(0, 2))
            .rot(OpKind.PHASE, 0, "p", controls=(1,)).CNOT(1, 2).RX(0, "b").build())
    obs = PauliSum(n, [PauliTerm(1.0, (), ((2, "Z"),)), PauliTerm(0.5, (), ((0, "X"), (1, "Z")))])
    rng = np.random.default_rng(2)
    names = prog.param_names
    v0 = rng.uniform(0.3, 2.5, size=(len(names), B))
    cp, cob = engine.CompiledProgram(prog), engine.CompiledObservable(obs)
    assert [cp.shift_rule(cp.slot_of[nm]) for nm in ("a", "b", "c", "d", "p")] == [2, 2, 4, 4, 2]

    def hess_diag_and_mixed(expect, v):
        e = expect(v)
        (g,) = torch.autograd.grad(e.sum(), v, create_graph=True)
        (h,) = torch.autograd.grad((g * torch.arange(1, len(names) + 1, device=g.device, dtype=g.dtype)[:, None]).sum(), v)
        return g.detach().cpu(), h.cpu()

    vh = torch.tensor(v0, requires_grad=True)
    want = hess_diag_and_mixed(lambda v: oracle.expectation(prog, [obs], {nm: v[i] for i, nm in enumerate(names)}), vh)
    vd = torch.tensor(v0, device=DEV, requires_grad=True)
    got = hess_diag_and_mixed(lambda v: engine.expectation(cp, [cob], {nm: v[i] for i, nm in enumerate(names)}, device=DEV), vd)
    assert (got[0] - want[0]).abs().max() < 1e-10
    assert (got[1] - want[1]).abs().max() < 1e-9


def test_shared_parameter_has_no_second_order_rule() -> None:
    from qadence_b200 import engine

    prog = ProgramBuilder(2).RX(0, "x").RX(1, "x").CNOT(0, 1).build()
    cp, cob = engine.CompiledProgram(prog), engine.CompiledObservable(total_magnetization(2))
    x = torch.tensor([0.3], dtype=torch.float64, device=DEV, requires_grad=True)
    e = engine.expectation(cp, [cob], {"x": x}, device=DEV)
    (g,) = torch.autograd.grad(e.sum(), x, create_graph=True)
    with pytest.raises(NotImplementedError):
        torch.autograd.grad(g.sum(), x)
