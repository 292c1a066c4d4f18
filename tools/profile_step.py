#!/usr/bin/env python
"""Two expectation+adjoint steps of cfg2 (hea(20, 8), batch 256, complex128) with nothing else on the device: the program
the ncu launch list / full captures under profiles/ are taken from (`ncu ... python tools/profile_step.py`).
Optional: --steps N, --dtype c64."""
from __future__ import annotations

import argparse
import ctypes as C
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import torch

from qadence_b200 import cabi, engine
from qadence_b200.engine import _DT, _stream_ptr
from qadence_b200.program import hea_program, total_magnetization


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--dtype", default="c128")
    ap.add_argument("--n", type=int, default=20)
    ap.add_argument("--depth", type=int, default=8)
    ap.add_argument("--batch", type=int, default=256)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    dtype = torch.complex128 if a.dtype == "c128" else torch.complex64
    lib = cabi.load()
    prog = hea_program(a.n, a.depth)
    cp = engine.CompiledProgram(prog)
    cob = engine.CompiledObservable(total_magnetization(a.n))
    cp, (cob,) = engine.absorb_trailing_permutation(cp, [cob])  # what engine.expectation runs (B200SV_ABSORB_PERM=0: the full circuit)
    seg = cp.segments[0]
    fwd = cp.device_plan(seg, dtype, False, dev)
    adj = cp.device_plan(seg, dtype, True, dev)
    coef = cob.pauli.coef({}, dev)
    g = torch.Generator().manual_seed(1234)
    ptable = (2 * torch.pi * torch.rand(len(cp.names), a.batch, generator=g, dtype=torch.float64)).to(dev)
    psi = torch.empty(a.batch, 1 << a.n, dtype=dtype, device=dev)
    lam = torch.empty_like(psi)
    grads = torch.zeros(len(cp.names), a.batch, dtype=torch.float64, device=dev)
    ws_f = fwd.workspace(dtype, a.batch, False).data_ptr()
    ws_a = adj.workspace(dtype, a.batch, True).data_ptr()
    st = _stream_ptr(dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    for _ in range(a.steps):
        cabi.check(lib.b200sv_init_zero_state(psi.data_ptr(), _DT[dtype], a.n, a.batch, st), "init")
        ev[0].record()
        cabi.check(lib.b200sv_apply_plan(psi.data_ptr(), _DT[dtype], a.n, a.batch, ptable.data_ptr(), C.byref(fwd.struct), 0, fwd.n_sweeps, 0, ws_f, st), "fwd")
        ev[1].record()
        e = cob.expectation(psi, {})
        cob.apply(psi, {}, out=lam)
        grads.zero_()
        ev[2].record()
        cabi.check(lib.b200sv_adjoint_plan(psi.data_ptr(), lam.data_ptr(), _DT[dtype], a.n, a.batch, ptable.data_ptr(), grads.data_ptr(), C.byref(adj.struct), 0,
                                           adj.n_sweeps, 0, ws_a, st), "adj")
        ev[3].record()
    torch.cuda.synchronize(dev)
    # CUDA-event times of the LAST step (the first one warms up): forward plan, Pauli kernels, adjoint plan
    print(f"TIMES forward_ms={ev[0].elapsed_time(ev[1]):.3f} ({fwd.n_sweeps} sweeps) pauli_ms={ev[1].elapsed_time(ev[2]):.3f} "
          f"adjoint_ms={ev[2].elapsed_time(ev[3]):.3f} ({adj.n_sweeps} sweeps) total_ms={ev[0].elapsed_time(ev[3]):.3f}")
    print("E[0:3] =", e[:3].tolist(), " |grad| =", float(grads.abs().sum()), " psi[0,0] =", complex(psi[0, 0]))


if __name__ == "__main__":
    main()
