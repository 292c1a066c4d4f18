"""configs[4]: chain(qft(n), hea(n, depth)) on a register sharded by its highest-order qubits.

    torchrun --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 tools/sharded_check.py --qubits 28 [--verify]

--verify: rank 0 also runs the same program on ONE GPU (n <= 29) and every rank compares its shard.
Always checks norm == 1; prints gates/s, HBM fraction of the local sweeps and NVLink bytes.
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
import torch.distributed as dist

from qadence_b200 import dist as qdist
from qadence_b200 import engine
from qadence_b200.program import Program, ProgramBuilder, hea_program, total_magnetization


def qft_program(n):
    b = ProgramBuilder(n)
    for q in range(n):
        b.H(q)
        for k in range(q + 1, n):
            b.CPHASE(k, q, 2 * np.pi / 2 ** (k - q + 1))
    return b.build()


ap = argparse.ArgumentParser()
ap.add_argument("--qubits", type=int, default=26)
ap.add_argument("--depth", type=int, default=1)
ap.add_argument("--verify", action="store_true")
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--peer", action="store_true", help="fuse the exchanges into the sweeps (NVLink peer-memory addressing)")
ap.add_argument("--grad", action="store_true", help="(with --peer) also expectation of a general Pauli sum + adjoint gradients of every parameter on the sharded register; with --verify compared with the single-GPU engine")
args = ap.parse_args()
rank, world = qdist.init_from_env("nccl")
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
n = args.qubits
prog = Program(n, qft_program(n).ops + hea_program(n, args.depth).ops)
g = torch.Generator().manual_seed(7)
values = {nm: 2 * torch.pi * torch.rand(1, generator=g, dtype=torch.float64) for nm in prog.param_names}
sim = qdist.PeerShardedSimulator(prog, rank, world, dev) if args.peer else qdist.QubitShardedSimulator(prog, rank, world, dev)
st = sim.run(values)  # warm-up (plans, NCCL communicator)
torch.cuda.synchronize()
times = []
for _ in range(args.reps):
    if not args.peer:
        del st  # a 36-qubit shard is 128 GiB: never hold two
        torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    st = sim.run(values)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    times.append(float(t.item()))
ms = min(times)
sweep_ms = None
if args.peer:
    sim.time_sweeps = True
    sim.run(values)
    sweep_ms = [round(x, 3) for x in sim.sweep_ms]
norm = float(sim.norm2(st).item())
mag = None if args.peer else float(sim.expectation_diagonal(st, total_magnetization(n)).item())
err = None
if args.verify:
    full = engine.run(engine.CompiledProgram(prog), values, device=dev) if rank == 0 else torch.empty(1, 1 << n, dtype=torch.complex128, device=dev)
    if world > 1:
        dist.broadcast(full, src=0)
    lo = rank << sim.n_local
    e = (st[0] - full[0, lo : lo + (1 << sim.n_local)]).abs().max()
    if world > 1:
        dist.all_reduce(e, op=dist.ReduceOp.MAX)
    err = float(e.item())
grad_info = None
if args.grad and args.peer:
    from qadence_b200.program import PauliSum, PauliTerm

    obs = PauliSum(n, [PauliTerm(1.0, (), ((q, "Z"),)) for q in range(n)] + [PauliTerm(0.7, (), ((0, "X"), (n - 1, "Y"))), PauliTerm(-0.4, (), ((1, "Y"), (n // 2, "Z")))])
    e_sh, g_sh = sim.expectation_and_grads(values, obs)  # warm-up (lambda shards, adjoint plan)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e_sh, g_sh = sim.expectation_and_grads(values, obs)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    grad_info = {"expectation": float(e_sh.item()), "n_params": int(g_sh.numel()), "ms_expectation_and_grads": float(t.item()),
                 "adjoint_sweeps": sim._plan_adj.n_sweeps}
    if args.verify:
        err_e = err_g = 0.0
        if rank == 0:
            leaf = {k: v.clone().requires_grad_(True) for k, v in values.items()}
            ref = engine.expectation(engine.CompiledProgram(prog), [engine.CompiledObservable(obs)], leaf, device=dev)
            ref.sum().backward()
            ref_g = torch.stack([leaf[nm].grad.reshape(()) for nm in prog.param_names]).to(dev)
            err_e = abs(float(ref.item()) - float(e_sh.item()))
            err_g = float((ref_g - g_sh).abs().max().item())
        grad_info.update({"abs_err_expectation_vs_single_gpu": err_e, "max_abs_err_gradient_vs_single_gpu": err_g})
if rank == 0:
    S_local = (1 << sim.n_local) * 16
    n_sweeps = sim.n_sweeps()
    print(json.dumps({
        "n": n, "world": world, "gates": len(prog.ops), "ms": ms, "gates_per_s": len(prog.ops) / (ms * 1e-3), "sweeps": n_sweeps,
        "mode": "peer-memory sweeps" if args.peer else "all-to-all exchanges",
        "exchanges": sim.stats["exchanges"], "nvlink_bytes_sent_per_rank": sim.stats["nvlink_bytes_per_rank_per_direction"] if args.peer else sim.bytes_sent,
        "global_bits_per_sweep": sim.stats.get("global_bits_per_sweep"), "sweep_ms_rank0": sweep_ms,
        "local_sweep_bytes": 2 * S_local * n_sweeps, "norm": norm, "total_magnetization": mag, "max_abs_err_vs_single_gpu": err, "grad": grad_info,
    }))
if args.peer:
    del st
    sim.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
