"""Randomised differential test of the device-side parameter embedding: random sympy expression trees over the
reference's function set → compile → CUDA forward / backward vs the torch restatement (oracle/embedding.py).

    python tools/gpu_fuzz_embedding.py [seconds] [seed]
"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import sympy
import torch

from oracle import embedding as oracle_embedding
from qadence_b200 import embedding as E

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
VARS = [sympy.Symbol(f"w{i}") for i in range(3)]
FEATS = [sympy.Symbol(f"x{i}") for i in range(2)]
UNARY = [sympy.sin, sympy.cos, sympy.tanh, sympy.exp, sympy.atan, sympy.Abs, sympy.Heaviside, sympy.asin, sympy.acos, sympy.log, sympy.tan, sympy.sqrt]


def rand_expr(depth: int):
    if depth == 0 or rng.random() < 0.25:
        k = rng.random()
        if k < 0.4:
            return VARS[int(rng.integers(3))]
        if k < 0.8:
            return FEATS[int(rng.integers(2))]
        return [sympy.Float(round(float(rng.normal()), 3)), sympy.Integer(int(rng.integers(1, 4))), sympy.Rational(1, 3), sympy.pi][int(rng.integers(4))]
    k = int(rng.integers(0, 5))
    if k == 0:
        return sympy.Add(*[rand_expr(depth - 1) for _ in range(int(rng.integers(2, 4)))])
    if k == 1:
        return sympy.Mul(*[rand_expr(depth - 1) for _ in range(int(rng.integers(2, 4)))])
    if k == 2:
        return sympy.Pow(rand_expr(depth - 1), [2, 3, sympy.Rational(1, 2), -1, VARS[0]][int(rng.integers(5))])
    f = UNARY[int(rng.integers(len(UNARY)))]
    return f(rand_expr(depth - 1))


t0 = time.time()
n_ok = n_skip = 0
while time.time() - t0 < budget:
    named = []
    try:
        for i in range(int(rng.integers(1, 12))):
            named.append((f"s{i}", rand_expr(int(rng.integers(0, 4)))))
        prog = E.compile_expressions(named, lambda nm: nm.startswith("x"))
    except (E.UnsupportedExpression, ValueError, TypeError, ZeroDivisionError):  # incl. sympy refusing e.g. Heaviside(acos(pi))
        n_skip += 1
        continue
    B = int(rng.integers(1, 70))
    var = torch.tensor(rng.uniform(0.2, 0.9, size=max(1, len(prog.var_names))), requires_grad=True)
    feat = torch.tensor(rng.uniform(0.05, 0.95, size=(max(1, len(prog.feat_names)), B)), requires_grad=True)
    want = oracle_embedding.evaluate(prog, var, feat, B)
    gsel = torch.tensor(rng.normal(size=tuple(want.shape)))
    finite = torch.isfinite(want.detach()).all(dim=1)  # rows that left a function's domain are compared as "both non-finite"
    if want.requires_grad:
        (torch.where(finite[:, None], want, torch.zeros_like(want)) * gsel).sum().backward()
    vd = var.detach().cuda().requires_grad_(True)
    fd = feat.detach().cuda().requires_grad_(True)
    got = E.evaluate(prog, vd, fd, B)
    g, w = got.detach().cpu(), want.detach()
    fin = torch.isfinite(w)
    if not torch.equal(torch.isfinite(g), fin):
        print("FORWARD MISMATCH (finiteness)", named)
        sys.exit(1)
    if fin.any() and (g[fin] - w[fin]).abs().max() > 1e-9 * max(1.0, float(w[fin].abs().max())):
        print("FORWARD MISMATCH", named, (g[fin] - w[fin]).abs().max().item())
        sys.exit(1)
    if not (want.requires_grad and got.requires_grad):
        n_ok += 1
        continue
    (torch.where(finite[:, None].cuda(), got, torch.zeros_like(got)) * gsel.cuda()).sum().backward()
    for a, b, what in ((vd.grad, var.grad, "var"), (fd.grad, feat.grad, "feat")):
        if a is None:
            continue
        a, b = a.cpu(), (b if b is not None else torch.zeros_like(a.cpu()))
        ok = torch.isfinite(b)
        if ok.any() and (a[ok] - b[ok]).abs().max() > 1e-7 * max(1.0, float(b[ok].abs().max())):
            print("BACKWARD MISMATCH", what, named, (a[ok] - b[ok]).abs().max().item())
            sys.exit(1)
    n_ok += 1
print(f"ok: {n_ok} programs ({n_skip} outside the device function set / node limit)")
