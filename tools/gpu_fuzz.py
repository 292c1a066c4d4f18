"""Randomised differential test of the CUDA sweeps against the CPU oracle: random circuits (controlled rotations / phases,
fixed gates, SWAP / CSWAP, random 2x2, Scale), random tile geometries, forward states and adjoint gradients.

    python tools/gpu_fuzz.py [seconds] [seed]
"""
import ctypes as C
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch

from oracle import statevec as oracle
from qadence_b200 import cabi, engine
from qadence_b200.engine import DevicePlan, _DT, _stream_ptr
from qadence_b200.plan import TileConfig, build_plan
from qadence_b200.program import OpKind, PauliSum, PauliTerm, ProgramBuilder
from tests.test_plan_cpu import random_program

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
dev = torch.device("cuda:0")
lib = cabi.load()
t0 = time.time()
n_fwd = n_adj = 0
worst_f = worst_g = 0.0
while time.time() - t0 < budget:
    n = int(rng.integers(1, 13))
    m = int(rng.integers(2, 13))
    r = int(rng.integers(1, min(4, m) + 1))
    if m - r > 8:
        continue
    low = int(rng.integers(0, min(m - r, 4) + 1))
    cfg = TileConfig(m=m, r=r, low=low, fold=int(rng.integers(2, 5)))
    B = int(rng.integers(1, 4))
    dtype = torch.complex128 if rng.random() < 0.8 else torch.complex64
    tol = 1e-10 if dtype == torch.complex128 else 2e-4
    # ---- forward: any gate set ----
    prog, names = random_program(n, int(rng.integers(1, 60)), rng)
    pn = prog.param_names
    values = {nm: torch.tensor(rng.uniform(-3, 3, size=B)) for nm in names}
    psi0 = torch.tensor(rng.normal(size=(B, 2**n)) + 1j * rng.normal(size=(B, 2**n)))
    psi0 = psi0 / psi0.abs().pow(2).sum(1, keepdim=True).sqrt()
    want = oracle.run(prog, values, psi0)
    dp = DevicePlan(build_plan(prog.ops, n, {nm: i for i, nm in enumerate(pn)}, cfg), dev)
    tab = engine.pack_params(pn, values, B, dev)
    st = psi0.to(device=dev, dtype=dtype).clone()
    rc = lib.b200sv_apply_plan(st.data_ptr(), _DT[dtype], n, B, tab.data_ptr(), C.byref(dp.struct), 0, dp.n_sweeps, 0,
                               dp.workspace(dtype, B, False).data_ptr(), _stream_ptr(dev))
    assert rc == 0, rc
    err = (st.cpu().to(torch.complex128) - want).abs().max().item() / max(1.0, want.abs().max().item())
    worst_f = max(worst_f, err if dtype == torch.complex128 else 0.0)
    if err > tol:
        print("FORWARD MISMATCH", n, cfg, dtype, err)
        print(prog.dumps())
        sys.exit(1)
    n_fwd += 1
    # ---- adjoint: unitary gates only (rotations, phases, fixed unitaries, SWAPs), controlled or not ----
    b = ProgramBuilder(n)
    anames = [f"a{i}" for i in range(5)]
    for _ in range(int(rng.integers(1, 40))):
        qs = [int(q) for q in rng.permutation(n)]
        nc = int(rng.integers(0, min(3, n))) if n > 1 else 0
        ctrl = qs[1 : 1 + nc]
        k = int(rng.integers(0, 7))
        p = anames[int(rng.integers(5))] if rng.random() < 0.8 else float(rng.normal())
        if k < 4:
            b.rot([OpKind.RX, OpKind.RY, OpKind.RZ, OpKind.PHASE][k], qs[0], p, ctrl)
        elif k == 4:
            b.gate(["X", "Y", "Z", "H", "S", "T"][int(rng.integers(6))], qs[0], ctrl)
        elif k == 5 and n >= 2:
            b.SWAP(qs[0], qs[1], qs[2 : 2 + min(nc, n - 2)])
        else:
            b.CNOT(qs[1], qs[0]) if n >= 2 else b.H(qs[0])
    aprog = b.build()
    apn = aprog.param_names
    terms = [PauliTerm(float(rng.normal()), (), tuple(sorted((int(q), "XYZ"[int(rng.integers(3))]) for q in rng.choice(n, size=int(rng.integers(1, min(n, 3) + 1)), replace=False))))
             for _ in range(int(rng.integers(1, 4)))]
    obs = PauliSum(n, terms)
    avals = {nm: torch.tensor(rng.uniform(-3, 3, size=B)) for nm in apn}
    e_want, g_want = oracle.adjoint_expectation_and_grads(aprog, obs, avals)
    cp = engine.CompiledProgram(aprog)
    am, ar, alow = [(10, 3, 2), (9, 3, 4), (8, 2, 2), (6, 1, 1), (10, 3, 0), (7, 3, 3), (5, 2, 0)][int(rng.integers(7))]
    seg_cfg = TileConfig(m=am, r=ar, low=alow, fold=cfg.fold)
    orig = TileConfig.default
    TileConfig.default = staticmethod(lambda c128, adj, _c=cfg, _s=seg_cfg: _s if adj else _c)  # type: ignore[assignment]
    try:
        dvals = {nm: v.clone().to(dev).requires_grad_(True) for nm, v in avals.items()}
        e = engine.expectation(cp, [engine.CompiledObservable(obs)], dvals, dtype=dtype, device=dev)
        if dvals:
            e.sum().backward()
    finally:
        TileConfig.default = orig  # type: ignore[assignment]
    etol = 1e-9 if dtype == torch.complex128 else 5e-4
    ee = (e.detach().cpu().reshape(-1).to(torch.float64) - e_want.reshape(-1)).abs().max().item()
    ge = max([(dvals[nm].grad.cpu() - g_want[nm].reshape(-1)).abs().max().item() for nm in apn if dvals[nm].grad is not None] or [0.0])
    worst_g = max(worst_g, ge if dtype == torch.complex128 else 0.0)
    if ee > etol or ge > etol * 10:
        print("ADJOINT MISMATCH", n, cfg, seg_cfg, dtype, ee, ge)
        print(aprog.dumps())
        sys.exit(1)
    n_adj += 1
print(f"ok: {n_fwd} forward cases (worst c128 rel err {worst_f:.2e}), {n_adj} adjoint cases (worst c128 grad err {worst_g:.2e})")
