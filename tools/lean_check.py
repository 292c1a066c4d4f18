#!/usr/bin/env python
"""Diagnostic for lean_sweep_kernel: run a small plan sweep by sweep on the GPU and compare the state after every sweep
with the CPU emulation of the same tables (tests/lean_emulator.py).  Prints one line per sweep; exits 1 on mismatch.

    python tools/lean_check.py [--n 13] [--depth 2] [--batch 2] [--c64]
"""
from __future__ import annotations

import argparse
import ctypes as C
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import numpy as np
import torch

from oracle import statevec as oracle
from qadence_b200 import cabi, engine
from qadence_b200.engine import DevicePlan, _DT, _stream_ptr
from qadence_b200.plan import TileConfig, build_plan
from qadence_b200.program import PauliSum, PauliTerm, hea_program
from tests.lean_emulator import emulate_lean_sweep


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=13)
    ap.add_argument("--depth", type=int, default=2)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--c64", action="store_true")
    args = ap.parse_args()
    n, B = args.n, args.batch
    dtype = torch.complex64 if args.c64 else torch.complex128
    tol = 1e-4 if args.c64 else 1e-10
    dev = torch.device("cuda:0")
    lib = cabi.load()
    prog = hea_program(n, args.depth)
    rng = np.random.default_rng(0)
    values = {nm: torch.tensor(rng.uniform(0.1, 3.0, size=B)) for nm in prog.param_names}
    slot_of = {nm: i for i, nm in enumerate(prog.param_names)}
    params = np.stack([values[nm].numpy() for nm in prog.param_names])
    tab = engine.pack_params(prog.param_names, values, B, dev)
    bad = 0
    print(f"env TMA={os.environ.get('B200SV_TMA', '1')} NBUF={os.environ.get('B200SV_LEAN_NBUF', '1')} dtype={dtype}")
    # forward
    fplan = build_plan(prog.ops, n, slot_of, TileConfig.default(not args.c64, False))
    dp = DevicePlan(fplan, dev)
    st = engine.zero_state(n, B, dtype, dev)
    ref = oracle.zero_state(n, B).numpy().astype(np.complex128)
    ws = dp.workspace(dtype, B, False)
    for s in range(dp.n_sweeps):
        rc = lib.b200sv_apply_plan(st.data_ptr(), _DT[dtype], n, B, tab.data_ptr(), C.byref(dp.struct), s, s + 1, 0, ws.data_ptr(), _stream_ptr(dev))
        torch.cuda.synchronize()
        mode = int(fplan.sweeps[s]["io_mode"])
        if mode == 0:
            print(f"fwd sweep {s}: generic kernel (skipping emulator), rc={rc}")
            ref = st.cpu().numpy().astype(np.complex128)
            continue
        emulate_lean_sweep(fplan, s, ref, None, params, True)
        err = np.abs(st.cpu().numpy() - ref).max()
        print(f"fwd sweep {s}: rc={rc} io_mode={mode} tile={list(fplan.sweeps[s]['tile_bits'][:fplan.sweeps[s]['m']])} rounds={int(fplan.sweeps[s]['n_rounds'])} max|d|={err:.3e}")
        if rc != 0 or not err < tol:
            bad += 1
            d = np.abs(st.cpu().numpy() - ref)
            idx = np.argwhere(d > tol)[:8]
            print("   first mismatches (b, i):", idx.tolist())
            ref = st.cpu().numpy().astype(np.complex128)  # resynchronise to localise later errors
    want = oracle.run(prog, values).numpy()
    print("fwd vs oracle:", np.abs(st.cpu().numpy() - want).max())
    # adjoint
    obs = PauliSum(n, [PauliTerm(1.0, (), ((q, "Z"),)) for q in range(n)])
    want_e, want_g = oracle.adjoint_expectation_and_grads(prog, obs, values)
    psi_t = oracle.run(prog, values)
    lam_t = oracle.from_pyq_shape(oracle.paulisum_apply(obs, oracle.to_pyq_shape(psi_t, n), values))
    aplan = build_plan(prog.ops, n, slot_of, TileConfig.default(not args.c64, True))
    dpa = DevicePlan(aplan, dev)
    dpsi, dlam = psi_t.to(device=dev, dtype=dtype).clone(), lam_t.to(device=dev, dtype=dtype).clone()
    rpsi, rlam = psi_t.numpy().copy(), lam_t.numpy().copy()
    grads = torch.zeros(len(prog.param_names), B, dtype=torch.float64, device=dev)
    rgrads = np.zeros((len(prog.param_names), B))
    wsa = dpa.workspace(dtype, B, True)
    for s in reversed(range(dpa.n_sweeps)):
        rc = lib.b200sv_adjoint_plan(dpsi.data_ptr(), dlam.data_ptr(), _DT[dtype], n, B, tab.data_ptr(), grads.data_ptr(), C.byref(dpa.struct), s, s + 1, 0,
                                     wsa.data_ptr(), _stream_ptr(dev))
        torch.cuda.synchronize()
        mode = int(aplan.sweeps[s]["io_mode"])
        if mode == 0:
            print(f"adj sweep {s}: generic kernel, rc={rc}")
            rpsi, rlam = dpsi.cpu().numpy().astype(np.complex128), dlam.cpu().numpy().astype(np.complex128)
            rgrads = grads.cpu().numpy().copy()
            continue
        emulate_lean_sweep(aplan, s, rpsi, rlam, params, False, rgrads)
        e1 = np.abs(dpsi.cpu().numpy() - rpsi).max()
        e2 = np.abs(dlam.cpu().numpy() - rlam).max()
        e3 = np.abs(grads.cpu().numpy() - rgrads).max()
        print(f"adj sweep {s}: rc={rc} io_mode={mode} rounds={int(aplan.sweeps[s]['n_rounds'])} max|dpsi|={e1:.3e} |dlam|={e2:.3e} |dgrad|={e3:.3e}")
        if rc != 0 or not max(e1, e2) < tol or not e3 < 100 * tol:
            bad += 1
            rpsi, rlam = dpsi.cpu().numpy().astype(np.complex128), dlam.cpu().numpy().astype(np.complex128)
            rgrads = grads.cpu().numpy().copy()
    gerr = max(float((grads[i].cpu() - want_g[nm]).abs().max()) for nm, i in slot_of.items())
    print("adjoint gradients vs oracle:", gerr)
    print("LEAN CHECK", "FAILED" if bad else "OK")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
