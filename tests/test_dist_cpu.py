"""Multi-GPU host logic on CPU: the global-qubit layout planner (ranks emulated in-process with the
plan emulator), and the real collectives (`exchange_global_local`, gradient all-reduce, batch
sharding) under a world_size-2 gloo group."""

from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import statevec as oracle
from qadence_b200 import dist as qdist
from qadence_b200.plan import TileConfig, build_plan
from qadence_b200.program import OpKind, Program, ProgramBuilder, hea_program
from tests.plan_emulator import emulate

ROOT = Path(__file__).resolve().parents[1]


def qft_program(n: int) -> Program:
    """qadence.constructors.qft (no final swaps): H(q) then CPHASE(k -> q, 2π/2^(k-q+1)) for k > q."""
    b = ProgramBuilder(n)
    for q in range(n):
        b.H(q)
        for k in range(q + 1, n):
            b.CPHASE(k, q, 2 * np.pi / 2 ** (k - q + 1))
    return b.build()


def run_sharded_emulated(prog: Program, values: dict, g: int, cfg: TileConfig) -> tuple[np.ndarray, qdist.LayoutPlan]:
    n = prog.n_qubits
    P = 1 << g
    n_local = n - g
    lp = qdist.plan_layout(prog.ops, n, g, restore=True)
    slot_of = {nm: i for i, nm in enumerate(prog.param_names)}
    params = np.stack([np.atleast_1d(values[nm].numpy()) for nm in prog.param_names]) if prog.param_names else np.zeros((1, 1))
    shards = [np.zeros((1, 1 << n_local), dtype=np.complex128) for _ in range(P)]
    shards[0][0, 0] = 1.0
    for step in lp.steps:
        if step.kind == "a2a":
            chunk = (1 << n_local) // P
            old = [s.reshape(P, chunk).copy() for s in shards]
            for r in range(P):
                for s_ in range(P):
                    shards[r].reshape(P, chunk)[s_] = old[s_][r]
            continue
        plan = build_plan(step.ops, n_local, slot_of, cfg, n_total=n)
        for r in range(P):
            shards[r] = emulate(plan, shards[r], params, index_hi=r << n_local)
    return np.concatenate([s[0] for s in shards]), lp


@pytest.mark.parametrize("g", [1, 2])
def test_qft_hea_sharded_matches_single(g: int) -> None:
    """BASELINE.json configs[4] at small n: chain(qft(n), hea(n, 1)) sharded by its top g qubits."""
    n = 7
    prog = Program(n, qft_program(n).ops + hea_program(n, 1).ops)
    rng = np.random.default_rng(g)
    values = {nm: torch.tensor(rng.uniform(0, 2 * np.pi, size=1)) for nm in prog.param_names}
    got, lp = run_sharded_emulated(prog, values, g, TileConfig(m=4, r=2, low=2, fold=3))
    want = oracle.run(prog, values)[0].numpy()
    np.testing.assert_allclose(got, want, atol=1e-12)
    assert lp.final_layout == list(range(n))
    assert 1 <= lp.n_exchanges <= 8
    # QFT: every CPHASE is diagonal, so only the n Hadamards can force an exchange
    q_only = qdist.plan_layout(qft_program(n).ops, n, g, restore=False)
    assert q_only.n_exchanges <= 2


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_circuit_sharded(seed: int) -> None:
    from tests.test_plan_cpu import random_program

    rng = np.random.default_rng(seed)
    n, g = 6, 2
    prog, names = random_program(n, 40, rng)
    prog = Program(n, [op for op in prog.ops if not (op.kind == OpKind.SWAP and op.controls)])  # CSWAP: not affine, not sharded yet
    values = {nm: torch.tensor(rng.uniform(0.5, 1.5, size=1)) for nm in prog.param_names}
    got, lp = run_sharded_emulated(prog, values, g, TileConfig(m=3, r=2, low=1, fold=3))
    want = oracle.run(prog, values)[0].numpy()
    np.testing.assert_allclose(got, want, atol=1e-10 * max(1.0, np.abs(want).max()))


def test_global_diagonal_gates_need_no_exchange() -> None:
    n, g = 8, 3
    b = ProgramBuilder(n)
    for q in range(n):
        b.RZ(q, 0.3 * (q + 1)).PHASE(q, 0.1).Z(q).S(q).T(q)
    for q in range(1, n):
        b.CZ(0, q).CPHASE(1, q, 0.2).CRZ(2, q, 0.4) if q > 2 else b.CZ(q, 7)
    assert qdist.plan_layout(b.build().ops, n, g, restore=False).n_exchanges == 0


def test_shard_range_covers_batch() -> None:
    for batch, world in ((4096, 8), (10, 4), (3, 2), (1, 2)):
        spans = [qdist.shard_range(batch, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == batch
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


# ---------------------------------------------------------------------------------------------------
def _worker(rank: int, world: int, port: int, tmp: str) -> None:
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    sys.path.insert(0, str(ROOT))
    r, w = qdist.init_from_env("gloo")
    assert (r, w) == (rank, world)
    # C2: in-place chunked exchange == transpose of the [rank, top-local] axes
    n_local = 6
    full = torch.arange(world * (1 << n_local), dtype=torch.float64).to(torch.complex128).reshape(world, world, -1)
    mine = full[rank].clone().reshape(1, -1)
    sent = qdist.exchange_global_local(mine, world, staging_bytes=16 * 5 * world)  # forces several ragged pieces
    want = full[:, rank, :].reshape(1, -1)
    assert torch.equal(mine, want)
    assert sent == (world - 1) * (1 << n_local) // world * 16
    # batch sharding + C1 gradient all-reduce
    batch = 5
    x = torch.arange(batch, dtype=torch.float64)
    theta = torch.nn.Parameter(torch.tensor([0.5], dtype=torch.float64))
    vals = qdist.shard_batch({"x": x, "theta": theta}, batch, rank, world)
    loss = (vals["x"] * vals["theta"]).sum()
    loss.backward()
    qdist.allreduce_shared_gradients([theta])
    assert torch.allclose(theta.grad, torch.tensor([x.sum()]))
    out = qdist.gather_batch((vals["x"] * 2).reshape(-1, 1), batch)
    assert torch.equal(out[:, 0], x * 2)
    dist.barrier()
    dist.destroy_process_group()
    Path(tmp, f"ok{rank}").write_text("ok")


def test_collectives_world_size_2_gloo(tmp_path) -> None:
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_layout_planner_rejects_registers_too_small_for_an_exchange() -> None:
    """An exchange swaps all g rank bits with g local qubits the current gate does not need; with too few local qubits the
    planner must say so (it used to trip an assertion — found by a randomised campaign of 15 000 sharded circuits)."""
    b = ProgramBuilder(6)
    b.H(0).SWAP(0, 5)
    with pytest.raises(ValueError, match="at least"):
        qdist.plan_layout(b.build().ops, 6, 3)
    assert qdist.plan_layout(b.build().ops, 7, 3).final_layout == list(range(7))
