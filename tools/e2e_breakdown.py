#!/usr/bin/env python
"""Where the end-to-end (qadence plugin, host tensors) step of cfg2 spends its time beyond the device work: wall-clock of each
stage with a device synchronisation in between (so the stages do not overlap: the sum is an upper bound of the e2e step)."""
from __future__ import annotations

import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch

import bench

assert bench._qadence_frontend()
from qadence import QuantumCircuit, hea
from qadence import total_magnetization as q_total_magnetization
from qadence.backends.api import backend_factory
from qadence.blocks.utils import parameters
from qadence.transpile.block import set_trainable

N, D, B = bench.N_QUBITS, bench.DEPTH, bench.BATCH
dev = torch.device("cuda", 0)
block = hea(N, D)
set_trainable(block, value=False)
bk = backend_factory("b200sv", diff_mode="adjoint")
conv = bk.convert(QuantumCircuit(N, block), q_total_magnetization(N))
qnames = sorted({p.name for p in parameters(block) if not p.is_number}, key=lambda s: int(s.split("_")[-1]))
host = bench.host_table(0)
x = {nm: host[i].clone().requires_grad_(True) for i, nm in enumerate(qnames)}


def step(sync: bool):
    t = [time.perf_counter()]
    for v in x.values():
        v.grad = None
    t.append(time.perf_counter())
    emb = conv.embedding_fn(conv.params, x)
    if sync:
        torch.cuda.synchronize()
    t.append(time.perf_counter())
    out = bk.expectation(conv.circuit, conv.observable, emb)
    if sync:
        torch.cuda.synchronize()
    t.append(time.perf_counter())
    s = out.sum()
    if sync:
        torch.cuda.synchronize()
    t.append(time.perf_counter())
    s.backward()
    torch.cuda.synchronize()
    t.append(time.perf_counter())
    return [(b - a) * 1e3 for a, b in zip(t, t[1:])]


for _ in range(2):
    step(True)
names = ["clear grads", "embedding_fn (pack + H2D + embed kernel)", "Backend.expectation (forward sweeps + <O>, D2H of E)", "sum", "backward (lambda = O psi, adjoint sweeps, embed backward, D2H, 480 leaf grads)"]
acc = [0.0] * 5
for _ in range(5):
    for i, v in enumerate(step(True)):
        acc[i] += v / 5
for nm, v in zip(names, acc):
    print(f"{v:8.3f} ms  {nm}")
print(f"{sum(acc):8.3f} ms  total (synchronised stages)")
t0 = time.perf_counter()
for _ in range(5):
    step(False)
print(f"{(time.perf_counter() - t0) / 5 * 1e3:8.3f} ms  per step, unsynchronised")
# finer: the backward pass under the autograd profiler (host side)
from torch.profiler import ProfilerActivity, profile

with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step(False)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=60))
