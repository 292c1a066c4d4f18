"""Randomised differential test of whole programs through `engine.run` / `engine.expectation(...).backward()`:
runs of gates interleaved with HamEvo (diagonal and Krylov, parametric time), dense unitaries on k qubits and Scale;
random Pauli-sum observables with a parametric coefficient; batch broadcasting of states.  Gradients are compared with
torch autograd through the CPU oracle.

    python tools/gpu_fuzz_mixed.py [seconds] [seed]
"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch

from oracle import statevec as oracle
from qadence_b200 import engine
from qadence_b200.program import Op, OpKind, PauliSum, PauliTerm, Program, ProgramBuilder

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 40.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
dev = torch.device("cuda:0")
t0 = time.time()
n_ok = 0
worst_e = worst_g = 0.0


def rand_paulisum(n, names=(), diagonal=False, max_terms=6):
    terms = []
    for _ in range(int(rng.integers(1, max_terms + 1))):
        k = int(rng.integers(1, min(n, 3) + 1))
        qs = sorted(int(q) for q in rng.choice(n, size=k, replace=False))
        terms.append(PauliTerm(float(rng.normal()), tuple(nm for nm in names if rng.random() < 0.4),
                               tuple((q, "Z" if diagonal else "XYZ"[int(rng.integers(3))]) for q in qs)))
    return PauliSum(n, terms)


while time.time() - t0 < budget:
    n = int(rng.integers(2, 9))
    B = int(rng.integers(1, 4))
    names = [f"p{i}" for i in range(4)]
    b = ProgramBuilder(n)
    for _seg in range(int(rng.integers(1, 5))):
        for _ in range(int(rng.integers(0, 16))):
            qs = [int(q) for q in rng.permutation(n)]
            nc = int(rng.integers(0, min(3, n)))
            k = int(rng.integers(0, 6))
            p = names[int(rng.integers(4))] if rng.random() < 0.7 else float(rng.normal())
            if k < 4:
                b.rot([OpKind.RX, OpKind.RY, OpKind.RZ, OpKind.PHASE][k], qs[0], p, qs[1 : 1 + nc])
            elif k == 4:
                b.gate(["X", "Y", "Z", "H", "S", "T"][int(rng.integers(6))], qs[0], qs[1 : 1 + nc])
            else:
                b.SWAP(qs[0], qs[1], qs[2 : 2 + min(nc, n - 2)])
        kind = int(rng.integers(0, 4))
        if kind == 0:
            b.hamevo(rand_paulisum(n, names=("g",), diagonal=rng.random() < 0.5), "t" if rng.random() < 0.7 else float(rng.uniform(0.1, 2)))
        elif kind == 1:
            k = int(rng.integers(2, min(n, 3) + 1))
            q, _r = np.linalg.qr(rng.normal(size=(2**k, 2**k)) + 1j * rng.normal(size=(2**k, 2**k)))
            b._add(Op(OpKind.MATK, tuple(int(x) for x in rng.choice(n, size=k, replace=False)), (), matrix=q, label="U"))
        elif kind == 2:
            b.scale("s" if rng.random() < 0.5 else float(rng.uniform(0.5, 1.5)))
    prog = b.build()
    obs = [rand_paulisum(n, names=("w",)), rand_paulisum(n)]
    all_names = list(prog.param_names) + ["w"]
    host = {nm: torch.tensor(rng.uniform(0.3, 1.6, size=B if rng.random() < 0.6 else 1), requires_grad=True) for nm in all_names}
    psi0 = None
    if rng.random() < 0.5:
        sb = B if rng.random() < 0.5 else 1
        psi0 = torch.tensor(rng.normal(size=(sb, 2**n)) + 1j * rng.normal(size=(sb, 2**n)))
        psi0 = psi0 / psi0.abs().pow(2).sum(1, keepdim=True).sqrt()
    sel = torch.tensor(rng.normal(size=(max([B] + [v.numel() for v in host.values()] + ([psi0.shape[0]] if psi0 is not None else [])), 2)))
    want = oracle.expectation(prog, obs, host, psi0)
    if want.requires_grad:
        (want * sel[: want.shape[0]]).sum().backward()
    devv = {nm: v.detach().clone().to(dev).requires_grad_(True) for nm, v in host.items()}
    cp = engine.CompiledProgram(prog)
    got = engine.expectation(cp, [engine.CompiledObservable(o) for o in obs], devv, psi0, device=dev)
    if got.requires_grad:
        (got * sel[: got.shape[0]].to(dev)).sum().backward()
    scale = max(1.0, float(want.detach().abs().max()))
    ee = (got.detach().cpu() - want.detach()).abs().max().item() / scale
    ge = 0.0
    for nm in all_names:
        gw = host[nm].grad if host[nm].grad is not None else torch.zeros_like(host[nm])
        gg = devv[nm].grad.cpu() if devv[nm].grad is not None else torch.zeros_like(gw)
        ge = max(ge, (gg - gw).abs().max().item() / max(1.0, float(gw.abs().max())) / scale)
    worst_e, worst_g = max(worst_e, ee), max(worst_g, ge)
    if ee > 1e-9 or ge > 1e-8:
        print("MISMATCH", n, B, "E", ee, "grad", ge)
        print(prog.dumps())
        sys.exit(1)
    st = engine.run(cp, {k: v.detach() for k, v in devv.items()}, psi0, device=dev).cpu()
    ws = oracle.run(prog, {k: v.detach() for k, v in host.items()}, psi0)
    if (st - ws).abs().max() > 1e-9 * max(1.0, float(ws.abs().max())):
        print("RUN MISMATCH", n, B, (st - ws).abs().max().item())
        print(prog.dumps())
        sys.exit(1)
    n_ok += 1
print(f"ok: {n_ok} programs (worst relative errors: expectation {worst_e:.2e}, gradient {worst_g:.2e})")
