#!/usr/bin/env python
"""bench.py — BASELINE.json's headline config on B200: expectation + adjoint parameter gradients of a
20-qubit HEA (depth 8), batch 256 independent parameter sets, complex128 (`configs[1]`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one full expectation+gradient evaluation of the whole batch: |0..0> init, fused
forward sweeps, <Z_tot>, λ = Oψ, fused adjoint sweeps that also reduce all 480 gradients.
Prints ONE JSON line (rank 0).  Multi-GPU (torchrun): every rank evaluates its own 256 parameter
sets (parameter batches shard trivially: no data-path collective) → "scaling": "weak".
"""

from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_QUBITS, DEPTH, BATCH = 20, 8, 256
FP64_PEAK_TFLOPS = 33.9  # measured on this pool's B200 with tools/ubench/dmma.cu (16 independent DFMA chains per thread)
CPU_SAMPLE_BATCH = 4  # parameter sets per CPU step (≈ 10-20 s of host work)
WORKLOAD = ("cfg2: hea(n_qubits=20, depth=8) expectation of total_magnetization + adjoint gradients of 480 parameters, "
            "batch 256 per GPU, complex128")
METRIC = "expectation+grad evals/s (HEA n=20 depth=8, batch 256, complex128, adjoint)"
UNIT = "evals/s"


def _peaks() -> tuple[float, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines: list[str] = []

    def start(self) -> None:
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self) -> None:
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def _cpu_eval(batch: int, threads: int) -> tuple[float, float]:
    """One expectation+gradient evaluation of hea(20,8) on the CPU oracle (the restated pyqtorch path:
    einsum per gate + Python adjoint loop).  Returns (seconds, evals/s)."""
    import torch

    from oracle import statevec as oracle
    from qadence_b200.program import hea_program, total_magnetization

    torch.set_num_threads(threads)
    prog = hea_program(N_QUBITS, DEPTH)
    obs = total_magnetization(N_QUBITS)
    g = torch.Generator().manual_seed(0)
    values = {nm: 2 * torch.pi * torch.rand(batch, generator=g, dtype=torch.float64) for nm in prog.param_names}
    t0 = time.perf_counter()
    oracle.adjoint_expectation_and_grads(prog, obs, values)
    dt = time.perf_counter() - t0
    return dt, batch / dt


def run_reference(args) -> None:
    """Reference arm: the reference's CPU path (restated — pyqtorch is not installable here) on the
    host cores, same config / metric, each step a bounded sample (batch 1 of the 256)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    threads = os.cpu_count() or 1
    sample_batch = CPU_SAMPLE_BATCH
    for _ in range(min(args.warmup, 1)):
        _cpu_eval(sample_batch, threads)
    ts = [_cpu_eval(sample_batch, threads)[0] for _ in range(args.steps)]
    dt = sum(ts) / len(ts)
    val = sample_batch / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "reference_step": f"{sample_batch} of the 256 parameter sets per step on the host cores (per-eval throughput)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample_batch} of the 256 parameter sets per step, torch {torch.__version__} einsum-per-gate + adjoint loop (oracle/statevec.py; pyqtorch is not installable here)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200sv")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank == 0 or world == 1:
        ge.build()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()

    from qadence_b200 import cabi, engine
    from qadence_b200.engine import _DT, _stream_ptr
    from qadence_b200.program import hea_program, total_magnetization

    lib = cabi.load()
    prog = hea_program(N_QUBITS, DEPTH)
    cp = engine.CompiledProgram(prog)
    cob = engine.CompiledObservable(total_magnetization(N_QUBITS))
    dtype = torch.complex128
    n_slots = len(cp.names)
    g = torch.Generator().manual_seed(1234 + rank)
    host_params = (2 * torch.pi * torch.rand(n_slots, BATCH, generator=g, dtype=torch.float64)).pin_memory()
    S = BATCH * (1 << N_QUBITS) * 16  # bytes of one state

    seg = cp.segments[0]
    fwd = cp.device_plan(seg, dtype, False, dev)
    adj = cp.device_plan(seg, dtype, True, dev)
    coef = cob.pauli.coef({}, dev)
    launches_per_step = 1 + (1 + fwd.n_sweeps) + 1 + 1 + (1 + adj.n_sweeps + 1)  # init, resolve+fwd sweeps, <O>, λ=Oψ, resolve+adj sweeps+grad
    ws_f = fwd.workspace(dtype, BATCH, False).data_ptr()
    ws_a = adj.workspace(dtype, BATCH, True).data_ptr()

    # ---------------- device-resident step (inputs already in HBM) ----------------
    ptable = host_params.to(dev)
    psi = torch.empty(BATCH, 1 << N_QUBITS, dtype=dtype, device=dev)
    lam = torch.empty_like(psi)
    grads = torch.zeros(n_slots, BATCH, dtype=torch.float64, device=dev)
    ones = torch.ones(BATCH, dtype=torch.float64, device=dev)

    def device_step():
        st = _stream_ptr(dev)
        cabi.check(lib.b200sv_init_zero_state(psi.data_ptr(), _DT[dtype], N_QUBITS, BATCH, st), "init")
        cabi.check(lib.b200sv_apply_plan(psi.data_ptr(), _DT[dtype], N_QUBITS, BATCH, ptable.data_ptr(), C.byref(fwd.struct), 0, fwd.n_sweeps, 0, ws_f, st), "fwd")
        e = engine.pauli_expectation(cob.pauli, coef, psi)
        engine.pauli_apply(cob.pauli, coef, psi, out=lam)
        grads.zero_()
        cabi.check(lib.b200sv_adjoint_plan(psi.data_ptr(), lam.data_ptr(), _DT[dtype], N_QUBITS, BATCH, ptable.data_ptr(), grads.data_ptr(), C.byref(adj.struct), 0, adj.n_sweeps, 0, ws_a, st), "adj")
        return e

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def timed(fn, steps):
        sync_all()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(steps):
            fn()
        ev1.record()
        sync_all()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(args.warmup):
        device_step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms = timed(device_step, args.steps)
    ms_per_step = ms / args.steps
    value = world * BATCH / (ms_per_step * 1e-3)

    # ---------------- end-to-end through the public API, host buffers ----------------
    names = cp.names
    d2h_e = torch.empty(BATCH, 1, dtype=torch.float64).pin_memory()
    d2h_g = torch.empty(n_slots, BATCH, dtype=torch.float64).pin_memory()

    def e2e_step():
        table = host_params.to(dev, non_blocking=True).requires_grad_(True)  # H2D of this step's inputs
        values = {nm: table[i] for i, nm in enumerate(names)}
        out = engine.expectation(cp, [cob], values, dtype=dtype, device=dev)
        out.sum().backward()
        d2h_e.copy_(out.detach(), non_blocking=True)  # D2H of the results
        d2h_g.copy_(table.grad, non_blocking=True)

    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps) / args.steps
    e2e_value = world * BATCH / (ms_e2e * 1e-3)
    clocks = sampler.stop() if rank == 0 else {}

    # ---------------- dominant kernel: the adjoint sweep (4·S algorithmic bytes per launch) ----------------
    st = _stream_ptr(dev)
    device_step()
    cabi.check(lib.b200sv_init_zero_state(psi.data_ptr(), _DT[dtype], N_QUBITS, BATCH, st), "init")
    cabi.check(lib.b200sv_apply_plan(psi.data_ptr(), _DT[dtype], N_QUBITS, BATCH, ptable.data_ptr(), C.byref(fwd.struct), 0, fwd.n_sweeps, 0, ws_f, st), "fwd")
    engine.pauli_apply(cob.pauli, coef, psi, out=lam)
    torch.cuda.synchronize(dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(adj.n_sweeps + 1)]
    ev[0].record()
    for i, s in enumerate(range(adj.n_sweeps - 1, -1, -1)):
        cabi.check(lib.b200sv_adjoint_plan(psi.data_ptr(), lam.data_ptr(), _DT[dtype], N_QUBITS, BATCH, ptable.data_ptr(), grads.data_ptr(), C.byref(adj.struct), s, s + 1, 0, ws_a, st), "adj")
        ev[i + 1].record()
    torch.cuda.synchronize(dev)
    adj_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(adj.n_sweeps)]
    # forward sweeps (2·S each), for the record
    evf = [torch.cuda.Event(enable_timing=True) for _ in range(fwd.n_sweeps + 1)]
    evf[0].record()
    for s in range(fwd.n_sweeps):
        cabi.check(lib.b200sv_apply_plan(psi.data_ptr(), _DT[dtype], N_QUBITS, BATCH, ptable.data_ptr(), C.byref(fwd.struct), s, s + 1, 0, ws_f, st), "fwd")
        evf[s + 1].record()
    torch.cuda.synchronize(dev)
    fwd_ms = [evf[i].elapsed_time(evf[i + 1]) for i in range(fwd.n_sweeps)]
    # a light sweep (one round: three merged rotations) shows the memory path alone: the state is streamed
    # through shared memory exactly like in the heavy sweeps, with 3 instead of ~10 merged 2x2 per tile
    from qadence_b200.program import OpKind, ProgramBuilder

    lb = ProgramBuilder(N_QUBITS)
    for q in (0, 1, 2):
        lb.rot(OpKind.RX, q, cp.names[q]).rot(OpKind.RY, q, cp.names[q + 20])
    lcp = engine.CompiledProgram(lb.build())
    lplan = lcp.device_plan(lcp.segments[0], dtype, False, dev)
    ltab = ptable[: max(1, len(lcp.names))] if False else engine.pack_params(lcp.names, {nm: ptable[cp.slot_of[nm]] for nm in lcp.names}, BATCH, dev)
    lws = lplan.workspace(dtype, BATCH, False).data_ptr()
    for _ in range(2):
        cabi.check(lib.b200sv_apply_plan(psi.data_ptr(), _DT[dtype], N_QUBITS, BATCH, ltab.data_ptr(), C.byref(lplan.struct), 0, lplan.n_sweeps, 0, lws, st), "light")
    torch.cuda.synchronize(dev)
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0.record()
    for _ in range(5):
        cabi.check(lib.b200sv_apply_plan(psi.data_ptr(), _DT[dtype], N_QUBITS, BATCH, ltab.data_ptr(), C.byref(lplan.struct), 0, lplan.n_sweeps, 0, lws, st), "light")
    l1.record()
    torch.cuda.synchronize(dev)
    light_ms = l0.elapsed_time(l1) / 5 / max(1, lplan.n_sweeps)
    light_achieved = 2 * S / (light_ms * 1e-3) / 1e9
    peak, peak_src = _peaks()
    adj_avg = sum(adj_ms) / len(adj_ms)
    fwd_avg = sum(fwd_ms) / len(fwd_ms)
    achieved = 4 * S / (adj_avg * 1e-3) / 1e9
    fwd_achieved = 2 * S / (fwd_avg * 1e-3) / 1e9

    # FP64 pipe view of the same step: every merged 2x2 costs 8 FP64 instructions per amplitude going forward and
    # 24 going backward (un-apply on psi and lambda + the four gradient moments), see DESIGN.md §4.1
    n_chain_f = int((fwd.plan.rounds["mslot"] >= 0).sum()) + int((fwd.plan.rounds["mslot4"] >= 0).sum())
    n_chain_a = int((adj.plan.rounds["mslot"] >= 0).sum()) + int((adj.plan.rounds["mslot4"] >= 0).sum())
    fp64_instr = (6 * n_chain_f + 18 * n_chain_a) * float(BATCH) * float(1 << N_QUBITS)
    fp64_tflops = 2 * fp64_instr / (ms_per_step * 1e-3) / 1e12

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            dt, v = _cpu_eval(CPU_SAMPLE_BATCH, threads)
            cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"{CPU_SAMPLE_BATCH} of the 256 parameter sets (one full expectation+adjoint eval of that sub-batch, {dt:.1f} s), einsum-per-gate restatement of the pyqtorch path (oracle/statevec.py)"}
        n_gates = len(prog.ops)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 (complex128)", "data": "synthetic",
            "config": {
                "workload": WORKLOAD,
                "l2": "inputs larger than L2 (state 4 GiB per GPU)", "gates": n_gates, "params": n_slots,
                "fwd_sweeps": fwd.n_sweeps, "adj_sweeps": adj.n_sweeps,
                "gates_per_s": world * n_gates * BATCH / (ms_per_step * 1e-3),
                "parallelism": f"batch-sharded x{world}",
            },
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": n_slots * BATCH * 8, "d2h_bytes_per_step": BATCH * 8 + n_slots * BATCH * 8},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": {"bound": "hbm", "kernel": "sweep_kernel<double,3,true,false> (adjoint sweep: psi+lambda read+write, gradient moments reduced)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": 17.134e9, "traffic_source": "profiles/r01h_final_build.md: dram__bytes_read 8.592 GB + dram__bytes_write 8.541 GB for one adjoint sweep (ncu --set full)",
                         "peak_source": peak_src, "bytes_per_launch": 4 * S, "avg_launch_ms": adj_avg,
                         "forward_sweep": {"achieved": fwd_achieved, "frac": fwd_achieved / peak, "bytes_per_launch": 2 * S, "avg_launch_ms": fwd_avg},
                         "light_sweep": {"what": "same kernel, one round of three merged 2x2 (6 gates) per sweep", "achieved": light_achieved,
                                         "frac": light_achieved / peak, "bytes_per_launch": 2 * S, "avg_launch_ms": light_ms},
                         "fp64_pipe": {"what": "whole step: FP64 instructions of the merged 2x2 applications and gradient moments (counted from the plan) vs the DFMA rate measured by tools/ubench/dmma.cu",
                                       "merged_2x2_fwd": n_chain_f, "merged_2x2_adj": n_chain_a, "achieved": fp64_tflops, "peak": FP64_PEAK_TFLOPS,
                                       "unit": "TFLOP/s", "frac": fp64_tflops / FP64_PEAK_TFLOPS,
                                       "floor_ms_per_step": 2 * fp64_instr / (FP64_PEAK_TFLOPS * 1e12) * 1e3},
                         "step": {"what": "algorithmic HBM bytes of the whole step (2S per forward sweep, 4S per adjoint sweep, S init, S <O>, 2S lambda=O psi): the tile-search planner trades bandwidth fraction per sweep for FEWER sweeps, i.e. fewer bytes per step",
                                  "hbm_bytes_per_step": (2 * fwd.n_sweeps + 4 * adj.n_sweeps + 4) * S,
                                  "hbm_floor_ms_per_step": (2 * fwd.n_sweeps + 4 * adj.n_sweeps + 4) * S / (peak * 1e9) * 1e3,
                                  "rounds_fwd": int(fwd.plan.sweeps["n_rounds"].sum()), "rounds_adj": int(adj.plan.sweeps["n_rounds"].sum())},
                         "note": "complex128 sweeps with ~20 merged 2x2 per tile are FP64-pipe/issue bound on B200, not HBM bound (an FP64 instruction holds the issue port 2 cycles; DMMA runs at the same pipe rate), see DESIGN.md §4.1 and profiles/r01_fp64_microbench.md"},
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
