"""Device-side parameter embedding (reference: qadence/blocks/embedding.py:118-167).

CPU: the expression compiler + the torch restatement of the interpreter (oracle/embedding.py) against the reference's own
host embedding on the reference's feature maps and parameter expressions.  GPU: the CUDA kernels against the restatement."""

from __future__ import annotations

import numpy as np
import pytest
import sympy
import torch

from oracle import embedding as oracle_embedding
from qadence_b200 import embedding as E

x, y, w, v = sympy.symbols("x y w v")
EXPRS = [
    ("plain", w),
    ("scaled", 2.5 * w),
    ("cheb", 2 * sympy.acos(x)),
    ("tower", 3 * sympy.acos(x) + w),
    ("fourier", sympy.pi * x * y),
    ("poly", x**2 * w - y**3 / 7 + 1),
    ("trig", sympy.sin(x) * sympy.cos(y + w) + sympy.tan(x / 3)),
    ("hyper", sympy.tanh(w * x) + sympy.exp(-y) + sympy.log(x + 2)),
    ("inv", sympy.asin(x) - sympy.atan(y * v) + sympy.Abs(w - x)),
    ("heavi", sympy.Heaviside(x - 0.5) * w + sympy.sqrt(x + 1)),
    ("powsym", (x + 1.5) ** w),
    ("const", sympy.Rational(1, 3) + sympy.E),
    ("shared", w),  # the same root feeding a second slot
]
IS_FEATURE = lambda nm: nm in ("x", "y")  # noqa: E731


def _inputs(batch: int, seed: int = 0):
    rng = np.random.default_rng(seed)
    var = torch.tensor(rng.uniform(0.2, 1.2, size=2), requires_grad=True)  # w, v in order of first use
    feat = torch.tensor(rng.uniform(0.05, 0.9, size=(2, batch)), requires_grad=True)
    return var, feat


def _sympy_reference(prog: E.EmbedProgram, var: torch.Tensor, feat: torch.Tensor) -> np.ndarray:
    out = []
    for nm, e in EXPRS:
        syms = sorted(sympy.sympify(e).free_symbols, key=str)
        fn = sympy.lambdify(syms, e, modules="numpy")
        args = [feat[prog.feat_names.index(str(s))].detach().numpy() if IS_FEATURE(str(s)) else float(var[prog.var_names.index(str(s))]) for s in syms]
        out.append(np.broadcast_to(np.asarray(fn(*args), dtype=np.float64), (feat.shape[1],)))
    return np.stack(out)


def test_compiler_and_restatement_match_sympy() -> None:
    prog = E.compile_expressions(EXPRS, IS_FEATURE)
    assert prog.var_names == ["w", "v"] and prog.feat_names == ["x", "y"]
    assert prog.n_slots == len(EXPRS) and int(prog.slot_off[-1]) == len(prog.nodes)
    assert sorted(prog.nodes["load"][np.isin(prog.nodes["op"], (E.EO_VAR, E.EO_FEAT))].tolist()) == list(range(prog.n_loads))
    var, feat = _inputs(5)
    got = oracle_embedding.evaluate(prog, var, feat, 5)
    assert np.abs(got.detach().numpy() - _sympy_reference(prog, var, feat)).max() < 1e-13


def test_unsupported_expressions_are_reported() -> None:
    with pytest.raises(E.UnsupportedExpression):
        E.compile_expressions([("a", sympy.erf(x))], IS_FEATURE)
    with pytest.raises(E.UnsupportedExpression):
        E.compile_expressions([("a", sympy.I * x)], IS_FEATURE)
    long = sum(sympy.sin(x + i) * sympy.cos(y * i) for i in range(1, 12))
    with pytest.raises(E.UnsupportedExpression):
        E.compile_expressions([("a", long)], IS_FEATURE)


def test_param_table_behaves_like_the_embedded_dict() -> None:
    table = torch.arange(6, dtype=torch.float64).reshape(3, 2)
    pt = E.ParamTable(["a", "b", "c"], table, {"x": torch.tensor([1.0, 2.0])}, torch.device("cpu"))
    assert len(pt) == 4 and list(pt) == ["a", "b", "c", "x"] and "b" in pt and "z" not in pt
    assert torch.equal(pt["b"], table[1]) and torch.equal(pt.get("x"), torch.tensor([1.0, 2.0])) and pt.get("z") is None
    assert torch.equal(pt.rows(["c", "a"]), table[[2, 0]]) and pt.rows(["a", "b", "c"]) is table and pt.rows(["a", "x"]) is None
    plain = dict(pt)
    assert set(plain) == {"a", "b", "c", "x"} and torch.equal({**pt}["c"], table[2])
    assert pt.clean
    pt["a"] = torch.zeros(2)  # e.g. a parameter-shift rule editing one entry: the table is no longer authoritative
    assert not pt.clean and torch.equal(pt["a"], torch.zeros(2))
    with pytest.raises(KeyError):
        pt["nope"]


@pytest.mark.gpu
@pytest.mark.parametrize("batch", [1, 7, 300])
def test_gpu_kernels_match_restatement(batch: int) -> None:
    prog = E.compile_expressions(EXPRS, IS_FEATURE)
    var, feat = _inputs(batch, seed=batch)
    want = oracle_embedding.evaluate(prog, var, feat, batch)
    gsel = torch.tensor(np.random.default_rng(1).normal(size=tuple(want.shape)))
    (want * gsel).sum().backward()
    vd = var.detach().cuda().requires_grad_(True)
    fd = feat.detach().cuda().requires_grad_(True)
    got = E.evaluate(prog, vd, fd, batch)
    assert (got.detach().cpu() - want.detach()).abs().max() < 1e-12
    (got * gsel.cuda()).sum().backward()
    assert (vd.grad.cpu() - var.grad).abs().max() < 1e-10 * max(1.0, float(var.grad.abs().max()))
    assert (fd.grad.cpu() - feat.grad).abs().max() < 1e-10 * max(1.0, float(feat.grad.abs().max()))


@pytest.mark.gpu
def test_gpu_param_table_feeds_the_engine() -> None:
    """hea(4,1) behind a Chebyshev-like feature map: the table goes straight into the sweeps; values and gradients equal the
    host-packed path."""
    from oracle import statevec as oracle
    from qadence_b200 import engine
    from qadence_b200.program import ProgramBuilder, hea_program, total_magnetization

    n, B = 4, 6
    b = ProgramBuilder(n)
    for q in range(n):
        b.RX(q, f"fm{q}")
    prog = b.build()
    prog.ops += hea_program(n, 1).ops
    theta = [nm for nm in prog.param_names if not nm.startswith("fm")]
    syms = {nm: sympy.Symbol(nm) for nm in theta}
    named = [(f"fm{q}", (q + 1) * sympy.acos(x)) for q in range(n)] + [(nm, syms[nm]) for nm in theta]
    eprog = E.compile_expressions(named, lambda nm: nm == "x")
    rng = np.random.default_rng(0)
    var = torch.tensor(rng.uniform(0, 3, size=len(theta)), device="cuda", requires_grad=True)
    xs = torch.tensor(rng.uniform(-0.9, 0.9, size=(1, B)), device="cuda", requires_grad=True)
    table = E.evaluate(eprog, var, xs, B)
    pt = E.ParamTable(eprog.names, table, {}, torch.device("cpu"))
    cp = engine.CompiledProgram(prog)
    cob = engine.CompiledObservable(total_magnetization(n))
    e = engine.expectation(cp, [cob], pt, device="cuda:0")
    e.sum().backward()
    # the host-dict path through pack_params (parity-tested against the oracle elsewhere) must give the same numbers
    vh = var.detach().clone().requires_grad_(True)
    xh = xs.detach().clone().requires_grad_(True)
    vals = {f"fm{q}": (q + 1) * torch.acos(xh[0]) for q in range(n)}
    vals.update({nm: vh[eprog.var_names.index(nm)].reshape(1) for nm in theta})
    e2 = engine.expectation(cp, [cob], vals, device="cuda:0")
    e2.sum().backward()
    assert (e.detach() - e2.detach()).abs().max() < 1e-12
    assert (var.grad - vh.grad).abs().max() < 1e-10
    assert (xs.grad - xh.grad).abs().max() < 1e-10
    want = oracle.expectation(prog, [total_magnetization(n)], {k: t.detach().cpu() for k, t in vals.items()})
    assert (e.detach().cpu() - want).abs().max() < 1e-10


def test_packed_feature_rows_are_one_autograd_node_with_the_gradients_of_a_stack() -> None:
    """`embedding._PackRows` (host side of the plugin's input path): same values and gradients as `torch.stack(...)` for rows of
    every accepted shape — [batch], [1], 0-dim, [batch, 1], float32 — including rows that do not ask for a gradient."""
    from qadence_b200.embedding import _PackRows

    B = 5
    g = torch.Generator().manual_seed(3)
    rows = [torch.rand(B, generator=g, dtype=torch.float64).requires_grad_(True), torch.rand(1, generator=g, dtype=torch.float64).requires_grad_(True),
            torch.tensor(0.7, dtype=torch.float64, requires_grad=True), torch.rand(B, 1, generator=g, dtype=torch.float64).requires_grad_(True),
            torch.rand(B, generator=g, dtype=torch.float32).requires_grad_(True), torch.rand(B, generator=g, dtype=torch.float64)]
    w = torch.rand(len(rows), B, generator=g, dtype=torch.float64)
    out = _PackRows.apply(torch.device("cpu"), B, *rows)
    assert out.shape == (len(rows), B) and out.dtype == torch.float64 and out.grad_fn is not None
    (out * w).sum().backward()
    got = [None if r.grad is None else r.grad.clone() for r in rows]
    for r in rows:
        r.grad = None
    ref = torch.stack([r.reshape(-1).to(torch.float64).expand(B) for r in rows])
    assert torch.equal(ref.detach(), out.detach())
    (ref * w).sum().backward()
    for a, r in zip(got, rows):
        assert (a is None) == (r.grad is None)
        if a is not None:
            assert a.shape == r.shape and a.dtype == r.dtype and torch.allclose(a, r.grad, atol=1e-6 if r.dtype == torch.float32 else 1e-14)
    # the common case (all rows [batch] float64) takes the single-copy path and keeps ONE node for all inputs
    plain = [torch.rand(B, generator=g, dtype=torch.float64).requires_grad_(True) for _ in range(4)]
    o2 = _PackRows.apply(torch.device("cpu"), B, *plain)
    assert type(o2.grad_fn).__name__ == "_PackRowsBackward" and len(o2.grad_fn.next_functions) == len(plain)  # straight to the leaves
    o2.sum().backward()
    assert all(torch.equal(p.grad, torch.ones(B, dtype=torch.float64)) for p in plain)
