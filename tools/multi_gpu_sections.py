"""The multi-GPU sections `bench.py --gpus N` (N > 1) adds to its JSON line, outside `value`:

* `collective` — configs[3]'s per-GPU share: feature map + hea(16, 12), 512 samples per GPU, forward + adjoint, then the
  gradient all-reduce the reference does through DDP (`ml_tools/train_utils/accelerator.py:285-305`) as ONE all-reduce of the
  576 shared-parameter gradients on the adjoint's stream: ms per step with and without it.
* `sharded`    — configs[4]'s shape: qft(n) + hea(n, 1) on ONE register sharded by its top qubits over the N GPUs (peer-memory
  sweeps over NVLink, `PeerShardedSimulator`), n the largest that keeps the run short: seconds, NVLink GB/s of the straddling
  sweeps against the measured peer bandwidth, HBM fraction of the local sweeps, and max |Δ| against a single GPU at n <= 29.
  Also expectation + adjoint gradients on the sharded register (`expectation_and_grads`).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.distributed as dist

NVLINK_PEER_GBS = 770.0  # measured peer copy, per direction per GPU (B200_PROFILING.md)


def _max_ms(ms: float, dev) -> float:
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def collective_section(rank: int, world: int, dev, steps: int) -> dict:
    from qadence_b200 import dist as qdist
    from qadence_b200 import engine
    from qadence_b200.program import Program, ProgramBuilder, hea_program, total_magnetization

    n, depth, per_gpu = 16, 12, 512
    b = ProgramBuilder(n)
    for q in range(n):
        b.RX(q, "x")
    prog = Program(n, b.build().ops + hea_program(n, depth).ops)
    cp = engine.CompiledProgram(prog)
    cob = engine.CompiledObservable(total_magnetization(n))
    g = torch.Generator().manual_seed(11)  # the same θ on every rank (shared parameters), different samples
    theta = {nm: (2 * math.pi * torch.rand(1, generator=g, dtype=torch.float64)).to(dev).requires_grad_(True) for nm in prog.param_names if nm != "x"}
    gx = torch.Generator().manual_seed(100 + rank)
    x = (torch.rand(per_gpu, generator=gx, dtype=torch.float64) * 1.98 - 0.99).to(dev)
    params = list(theta.values())

    def step(allreduce: bool):
        for p in params:
            p.grad = None
        e = engine.expectation(cp, [cob], {**theta, "x": x}, device=dev)
        loss = (e[:, 0] ** 2).mean()
        loss.backward()
        if allreduce:
            qdist.allreduce_shared_gradients(params)
        return loss

    out = {}
    for name, ar in (("without_allreduce", False), ("with_allreduce", True)):
        for _ in range(2):
            step(ar)
        torch.cuda.synchronize(dev)
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step(ar)
        e1.record()
        torch.cuda.synchronize(dev)
        out[name] = _max_ms(e0.elapsed_time(e1) / steps, dev)
    g0 = torch.cat([p.grad.reshape(-1) for p in params])
    chk = g0.clone()
    dist.broadcast(chk, 0)
    same = float((chk - g0).abs().max().item())  # after the all-reduce every rank holds the same gradient
    return {"workload": f"cfg4 per-GPU share: feature map + hea({n}, {depth}), {per_gpu} samples per GPU (global batch {per_gpu * world}), loss = mean(E^2), forward + adjoint + all-reduce of {len(params)} shared gradients",
            "ms_per_step_without_allreduce": out["without_allreduce"], "ms_per_step_with_allreduce": out["with_allreduce"],
            "allreduce_ms": out["with_allreduce"] - out["without_allreduce"], "allreduce_bytes": len(params) * 8,
            "samples_per_s": per_gpu * world / (out["with_allreduce"] * 1e-3), "max_gradient_disagreement_across_ranks": same}


def sharded_section(rank: int, world: int, dev, peak_hbm: float) -> dict:
    from qadence_b200 import dist as qdist
    from qadence_b200 import engine
    from qadence_b200.program import PauliSum, PauliTerm, Program, ProgramBuilder, hea_program

    def qft_hea(n):
        b = ProgramBuilder(n)
        for q in range(n):
            b.H(q)
            for k in range(q + 1, n):
                b.CPHASE(k, q, 2 * np.pi / 2 ** (k - q + 1))
        return Program(n, b.build().ops + hea_program(n, 1).ops)

    g_bits = world.bit_length() - 1
    res = {}
    # (1) parity against one GPU at n = 28; (2) timing at the size where every GPU holds a 2^30-amplitude (16 GiB) shard
    for tag, n, verify in (("verify", 28, True), ("timing", 30 + g_bits, False)):
        prog = qft_hea(n)
        gen = torch.Generator().manual_seed(7)
        values = {nm: 2 * torch.pi * torch.rand(1, generator=gen, dtype=torch.float64) for nm in prog.param_names}
        sim = qdist.PeerShardedSimulator(prog, rank, world, dev)
        try:
            st = sim.run(values)
            torch.cuda.synchronize(dev)
            dist.barrier()
            sim.time_sweeps = True
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            st = sim.run(values)
            e1.record()
            torch.cuda.synchronize(dev)
            ms = _max_ms(e0.elapsed_time(e1), dev)
            sweep_ms = torch.tensor(sim.sweep_ms, dtype=torch.float64, device=dev)
            dist.all_reduce(sweep_ms, op=dist.ReduceOp.MAX)
            sweep_ms = sweep_ms.tolist()
            ks = sim.stats["global_bits_per_sweep"]
            shard = (1 << sim.n_local) * 16
            local = [m for m, k in zip(sweep_ms, ks) if k == 0]
            strad = [(m, k) for m, k in zip(sweep_ms, ks) if k > 0]
            entry = {"n_qubits": n, "gates": len(prog.ops), "seconds": ms * 1e-3, "sweeps": len(ks), "global_bits_per_sweep": ks,
                     "sweep_ms": [round(v, 3) for v in sweep_ms], "norm": float(sim.norm2(st).item()),
                     "local_sweeps_hbm_frac": (2 * shard / (np.mean(local) * 1e-3) / 1e9 / peak_hbm) if local else None,
                     "straddling_sweeps_nvlink_gbs_per_direction": [round(shard * (1 - 2.0 ** -k) / (m * 1e-3) / 1e9, 1) for m, k in strad],
                     "nvlink_peak_gbs": NVLINK_PEER_GBS}
            if verify:
                full = engine.run(engine.CompiledProgram(prog), values, device=dev) if rank == 0 else torch.empty(1, 1 << n, dtype=torch.complex128, device=dev)
                dist.broadcast(full, src=0)
                lo = rank << sim.n_local
                err = (st[0] - full[0, lo: lo + (1 << sim.n_local)]).abs().max()
                dist.all_reduce(err, op=dist.ReduceOp.MAX)
                entry["max_abs_err_vs_single_gpu"] = float(err.item())
                del full
                obs = PauliSum(n, [PauliTerm(1.0, (), ((q, "Z"),)) for q in range(n)] + [PauliTerm(0.7, (), ((0, "X"), (n - 1, "Y")))])
                e_sh, g_sh = sim.expectation_and_grads(values, obs)
                torch.cuda.synchronize(dev)
                dist.barrier()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                e_sh, g_sh = sim.expectation_and_grads(values, obs)
                a1.record()
                torch.cuda.synchronize(dev)
                gms = _max_ms(a0.elapsed_time(a1), dev)
                errs = torch.zeros(2, dtype=torch.float64, device=dev)
                if rank == 0:
                    leaf = {k: v.clone().requires_grad_(True) for k, v in values.items()}
                    ref = engine.expectation(engine.CompiledProgram(prog), [engine.CompiledObservable(obs)], leaf, device=dev)
                    ref.sum().backward()
                    ref_g = torch.stack([leaf[nm].grad.reshape(()) for nm in prog.param_names]).to(dev)
                    errs[0] = abs(float(ref.item()) - float(e_sh.item()))
                    errs[1] = (ref_g - g_sh).abs().max()
                dist.broadcast(errs, 0)
                entry["expectation_and_grads"] = {"ms": gms, "n_params": int(g_sh.numel()), "abs_err_expectation_vs_single_gpu": float(errs[0]),
                                                  "max_abs_err_gradient_vs_single_gpu": float(errs[1])}
            res[tag] = entry
            del st
        finally:
            sim.close()
        torch.cuda.empty_cache()
    return {"workload": "cfg5 shape: qft(n) + hea(n, 1), complex128, ONE register sharded by its top qubits over the GPUs; sweeps read / write remote tiles over NVLink peer memory (no exchange passes)", **res}


def run(rank: int, world: int, dev, steps: int = 2) -> dict:
    import json
    from pathlib import Path

    p = Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json"
    peak = float(json.loads(p.read_text())["hbm_gbs"]) if p.exists() else 6650.0
    out = {}
    try:
        out["collective"] = collective_section(rank, world, dev, steps)
    except Exception as ex:  # a section must never take the headline number down with it
        out["collective"] = {"error": f"{type(ex).__name__}: {ex}"}
    try:
        if world & (world - 1) == 0:
            out["sharded"] = sharded_section(rank, world, dev, peak)
    except Exception as ex:
        out["sharded"] = {"error": f"{type(ex).__name__}: {ex}"}
    return out
