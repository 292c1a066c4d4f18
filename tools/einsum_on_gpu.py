#!/usr/bin/env python
"""Informative baseline ("and, informatively, the same on GPU"): the einsum-per-gate restatement of the pyqtorch
path (oracle/statevec.py: one `torch.einsum` per gate, the three-pass adjoint loop) with its tensors ON the B200 — what stock
ATen kernels deliver for cfg2, next to the hand-written sweeps.  Not part of the product, not part of the bench value.

    python tools/einsum_on_gpu.py [--batch 8]
"""
import argparse
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--qubits", type=int, default=20)
ap.add_argument("--depth", type=int, default=8)
args = ap.parse_args()
torch.set_default_device("cuda")  # BEFORE the oracle is imported: its constant matrices are created on the device too
from oracle import statevec as oracle  # noqa: E402
from qadence_b200.program import hea_program, total_magnetization  # noqa: E402

prog = hea_program(args.qubits, args.depth)
obs = total_magnetization(args.qubits)
g = torch.Generator(device="cpu").manual_seed(1234)
table = (2 * torch.pi * torch.rand(len(prog.param_names), 256, generator=g, dtype=torch.float64, device="cpu")).cuda()
res = {}
for B in (args.batch, 2 * args.batch):
    values = {nm: table[i, :B].clone() for i, nm in enumerate(prog.param_names)}
    oracle.adjoint_expectation_and_grads(prog, obs, values)  # warm-up (cuBLAS / einsum plans)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e, gr = oracle.adjoint_expectation_and_grads(prog, obs, values)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    res[f"batch_{B}"] = {"seconds": dt, "evals_per_s": B / dt}
# parity of this run against the CPU evaluation of the same code (first element)
torch.set_default_device("cpu")
print(json.dumps({"what": "oracle/statevec.py (einsum per gate + adjoint loop) with CUDA tensors on one B200, hea(%d, %d), complex128" % (args.qubits, args.depth),
                  "expectation_first_element": float(e[0]), **res}))
