"""Peer-memory sweeps (`b200sv_apply_plan_peers`): a register sharded over GPUs by its top qubits, gates on global qubits
executed as sweeps whose tiles straddle shards (SURVEY.md §8e, configs[4]).  One GPU: the peer table has one entry and the
result must equal the ordinary path; two GPUs (when the box has them): two processes over CUDA IPC."""

from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from qadence_b200.program import Program, ProgramBuilder, hea_program

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def _program(n: int) -> Program:
    b = ProgramBuilder(n)
    for q in range(n):
        b.H(q)
        for k in range(q + 1, min(n, q + 6)):
            b.CPHASE(k, q, 2 * np.pi / 2 ** (k - q + 1))
    b.SWAP(0, n - 1).CNOT(n - 1, 0)
    return Program(n, b.build().ops + hea_program(n, 1).ops)


def test_single_rank_peer_table_equals_plain_path() -> None:
    from qadence_b200 import dist as qdist
    from qadence_b200 import engine

    n = 16
    prog = _program(n)
    g = torch.Generator().manual_seed(3)
    values = {nm: 6.28 * torch.rand(1, generator=g, dtype=torch.float64) for nm in prog.param_names}
    sim = qdist.PeerShardedSimulator(prog, 0, 1, "cuda:0")
    try:
        got = sim.run(values).clone()
        want = engine.run(engine.CompiledProgram(prog), values, device="cuda:0")
        assert (got - want).abs().max() < 1e-13  # same gates; the plain path may pick another tile geometry
        assert abs(float(sim.norm2(got)) - 1.0) < 1e-12
    finally:
        sim.close()


@pytest.mark.parametrize("world", [2, 4])
def test_rank_bit_addressing_with_peer_buffers_on_one_device(world: int) -> None:
    """`b200sv_apply_plan_peers` with `world` separately allocated shards that all live on THIS device: the rank-bit
    addressing (index bits >= n_local select the buffer), the per-rank tile slices and the straddling sweeps are
    exercised on a single-GPU box.  The "ranks" take turns on one stream, sweep by sweep (stream order is the barrier)."""
    import ctypes as C

    from qadence_b200 import cabi, engine
    from qadence_b200.engine import DevicePlan, _DT, _stream_ptr
    from qadence_b200.plan import TileConfig, build_plan

    n = 17
    g = world.bit_length() - 1
    n_local = n - g
    prog = _program(n)
    gen = torch.Generator().manual_seed(5)
    values = {nm: 6.28 * torch.rand(1, generator=gen, dtype=torch.float64) for nm in prog.param_names}
    dev = torch.device("cuda:0")
    want = engine.run(engine.CompiledProgram(prog), values, device=dev)
    slot_of = {nm: i for i, nm in enumerate(prog.param_names)}
    dp = DevicePlan(build_plan(list(prog.ops), n, slot_of, TileConfig.default(True, False), packer="greedy"), dev)
    shards = [torch.zeros(1, 1 << n_local, dtype=torch.complex128, device=dev) for _ in range(world)]
    shards[0][0, 0] = 1.0
    ptrs = (C.c_void_p * world)(*[s.data_ptr() for s in shards])
    tab = engine.pack_params(prog.param_names, values, 1, dev)
    ws = dp.workspace(torch.complex128, 1, False)
    lib = cabi.load()
    for s in range(dp.n_sweeps):
        for rank in range(world):
            rc = lib.b200sv_apply_plan_peers(ptrs, world, rank, _DT[torch.complex128], n, tab.data_ptr(), C.byref(dp.struct), s, s + 1,
                                             ws.data_ptr(), _stream_ptr(dev))
            assert rc == 0
    got = torch.cat([s.reshape(-1) for s in shards]).reshape(1, -1)
    assert (got - want).abs().max().item() < 1e-12


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_expectation_and_adjoint_gradients_on_one_device(world: int) -> None:
    """`b200sv_pauli_expectation_peers`, `b200sv_pauli_apply_peers` and `b200sv_adjoint_plan_peers` with `world` psi and
    lambda shards on THIS device (the "ranks" take turns on one stream): a general Pauli-sum observable with X / Y factors
    on GLOBAL qubits, and the adjoint gradients of every parameter, against the single-device engine path."""
    import ctypes as C

    from qadence_b200 import cabi, engine
    from qadence_b200.engine import DevicePlan, _DT, _stream_ptr
    from qadence_b200.plan import TileConfig, build_plan
    from qadence_b200.program import PauliSum, PauliTerm

    n = 16
    g = world.bit_length() - 1
    n_local = n - g
    prog = _program(n)
    gen = torch.Generator().manual_seed(9)
    values = {nm: (6.28 * torch.rand(1, generator=gen, dtype=torch.float64)).requires_grad_(True) for nm in prog.param_names}
    obs = PauliSum(n, [PauliTerm(1.0, (), ((q, "Z"),)) for q in range(n)]
                   + [PauliTerm(0.7, (), ((0, "X"), (n - 1, "Y"))), PauliTerm(-0.4, (), ((1, "Y"), (5, "Z"))), PauliTerm(0.3, (), ((0, "Z"), (1, "Z")))])
    dev = torch.device("cuda:0")
    cp = engine.CompiledProgram(prog)
    want = engine.expectation(cp, [engine.CompiledObservable(obs)], values, device=dev)
    want.sum().backward()
    want_g = torch.stack([values[nm].grad.reshape(()) for nm in prog.param_names])

    slot_of = {nm: i for i, nm in enumerate(prog.param_names)}
    fplan = DevicePlan(build_plan(list(prog.ops), n, slot_of, TileConfig(m=11, r=4, low=2, fold=3, lean=False), packer="greedy"), dev)
    aplan = DevicePlan(build_plan(list(prog.ops), n, slot_of, TileConfig(m=10, r=3, low=2, fold=3, adjoint=True, lean=False), packer="greedy"), dev)
    psi = [torch.zeros(1, 1 << n_local, dtype=torch.complex128, device=dev) for _ in range(world)]
    lam = [torch.zeros(1, 1 << n_local, dtype=torch.complex128, device=dev) for _ in range(world)]
    psi[0][0, 0] = 1.0
    pp = (C.c_void_p * world)(*[t.data_ptr() for t in psi])
    pl = (C.c_void_p * world)(*[t.data_ptr() for t in lam])
    detached = {k: v.detach() for k, v in values.items()}
    tab = engine.pack_params(prog.param_names, detached, 1, dev)
    lib = cabi.load()
    st = _stream_ptr(dev)
    dt = _DT[torch.complex128]
    wsf = fplan.workspace(torch.complex128, 1, False)
    for s in range(fplan.n_sweeps):
        for rank in range(world):
            assert lib.b200sv_apply_plan_peers(pp, world, rank, dt, n, tab.data_ptr(), C.byref(fplan.struct), s, s + 1, wsf.data_ptr(), st) == 0
    cps = engine.CompiledPauliSum.build(obs)
    coef = torch.view_as_real(cps.coef({}, dev).to(torch.complex128).contiguous())
    e = torch.zeros(1, dtype=torch.float64, device=dev)
    for rank in range(world):  # the all-reduce of the per-rank parts
        assert lib.b200sv_pauli_expectation_peers(pp, world, rank, dt, n, cps.terms(dev).data_ptr(), cps.n_terms, cps.n_single, coef.data_ptr(), e.data_ptr(), st) == 0
    assert abs(e.item() - want.item()) < 1e-11
    for rank in range(world):
        assert lib.b200sv_pauli_apply_peers(pp, world, rank, lam[rank].data_ptr(), dt, n, cps.terms(dev).data_ptr(), cps.n_terms, cps.n_single, coef.data_ptr(), 1.0, 0.0, st) == 0
    grads = [torch.zeros(len(prog.param_names), 1, dtype=torch.float64, device=dev) for _ in range(world)]
    wsa = [aplan.workspace(torch.complex128, 1, True).clone() for _ in range(world)]  # one workspace per "rank" (moment accumulators)
    for s in range(aplan.n_sweeps - 1, -1, -1):
        for rank in range(world):
            rc = lib.b200sv_adjoint_plan_peers(pp, pl, world, rank, dt, n, tab.data_ptr(), grads[rank].data_ptr(), C.byref(aplan.struct), s, s + 1, wsa[rank].data_ptr(), st)
            assert rc == 0
    got_g = sum(gr[:, 0] for gr in grads).cpu()
    assert (got_g - want_g).abs().max().item() < 1e-10
    back = torch.cat([t.reshape(-1) for t in psi])
    assert abs(back[0].item() - 1.0) < 1e-10 and back[1:].abs().max().item() < 1e-10  # psi walked back to |0...0>


@pytest.mark.parametrize("nproc", [1, 2])
def test_shard_qubits_configuration_through_the_plugin(nproc: int) -> None:
    """`Configuration(shard="qubits")`: QuantumModel expectation + gradients and sampling on a sharded register equal the
    single-GPU path (tools/sharded_plugin_check.py under torchrun; one process is enough to run the whole code path)."""
    if torch.cuda.device_count() < nproc:
        pytest.skip("needs more GPUs")
    import json

    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
           "--master-port", str(29551 + nproc), str(ROOT / "tools" / "sharded_plugin_check.py"), "--qubits", "16"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env={**os.environ})
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert out.returncode == 0 and line, out.stderr[-3000:]
    res = json.loads(line[-1])
    assert res["abs_err_expectation"] < 1e-10 and res["max_abs_err_gradient"] < 1e-10, res
    assert res["sampling"]["identical_indices"] >= res["sampling"]["shots"] - 2, res


def test_traffic_model_counts_global_bits() -> None:
    from qadence_b200.dist import peer_traffic
    from qadence_b200.plan import TileConfig, build_plan

    n, g = 24, 2
    prog = _program(n)
    plan = build_plan(list(prog.ops), n, {nm: i for i, nm in enumerate(prog.param_names)}, TileConfig.default(True, False))
    t = peer_traffic(plan, n, g)
    assert len(t["global_bits_per_sweep"]) == plan.n_sweeps and max(t["global_bits_per_sweep"]) >= 1
    assert 0 < t["nvlink_bytes_per_rank_per_direction"] < t["hbm_bytes_per_rank"]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_ranks_match_single_gpu() -> None:
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", str(ROOT / "tools" / "sharded_check.py"), "--qubits", "22", "--verify", "--peer", "--reps", "1"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env={**os.environ})
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert out.returncode == 0 and line, out.stderr[-2000:]
    import json

    res = json.loads(line[-1])
    assert res["max_abs_err_vs_single_gpu"] < 1e-13 and abs(res["norm"] - 1) < 1e-12
