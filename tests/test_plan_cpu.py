"""Planner (host logic) vs the oracle, on CPU: the packed sweep plan, interpreted record by record,
must reproduce the op-by-op einsum result for random circuits of every simple gate kind."""

from __future__ import annotations

import numpy as np
import pytest
import torch

from oracle import statevec as oracle
from qadence_b200.plan import TileConfig, build_plan
from qadence_b200.program import CONST_MATS, OpKind, ProgramBuilder, hea_program
from tests.plan_emulator import emulate, emulate_adjoint


def random_program(n: int, n_gates: int, rng: np.random.Generator, n_params: int = 6):
    b = ProgramBuilder(n)
    names = [f"p{i}" for i in range(n_params)]
    fixed = [k for k in CONST_MATS if k not in ("Zero",)]
    for _ in range(n_gates):
        kind = rng.integers(0, 8)
        qs = list(rng.permutation(n))
        t = int(qs[0])
        n_ctrl = int(rng.integers(0, min(3, n))) if n > 1 else 0
        ctrl = [int(q) for q in qs[1 : 1 + n_ctrl]]
        param = names[int(rng.integers(n_params))] if rng.random() < 0.7 else float(rng.normal())
        if kind == 0:
            b.gate(fixed[int(rng.integers(len(fixed)))], t, ctrl)
        elif kind == 1:
            b.rot(OpKind.RX, t, param, ctrl)
        elif kind == 2:
            b.rot(OpKind.RY, t, param, ctrl)
        elif kind == 3:
            b.rot(OpKind.RZ, t, param, ctrl)
        elif kind == 4:
            b.rot(OpKind.PHASE, t, param, ctrl)
        elif kind == 5 and n >= 2:
            b.SWAP(int(qs[0]), int(qs[1]), [int(q) for q in qs[2 : 2 + min(n_ctrl, n - 2)]])
        elif kind == 6:
            m = rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2))
            b._add(__import__("qadence_b200").Op(OpKind.MAT1, (t,), tuple(ctrl), matrix=m, label="rand"))
        else:
            b.scale(param if rng.random() < 0.5 else complex(rng.normal(), rng.normal()))
    return b.build(), names


CONFIGS = [
    TileConfig(m=11, r=3, low=4, fold=3),
    TileConfig(m=5, r=2, low=2, fold=3),
    TileConfig(m=6, r=4, low=3, fold=4),
    TileConfig(m=4, r=1, low=1, fold=3),
]


@pytest.mark.parametrize("n", [1, 2, 3, 5, 7])
@pytest.mark.parametrize("cfg", CONFIGS, ids=lambda c: f"m{c.m}r{c.r}")
def test_plan_matches_oracle(n: int, cfg: TileConfig) -> None:
    rng = np.random.default_rng(100 * n + cfg.m)
    prog, names = random_program(n, 40, rng)
    B = 3
    values = {nm: torch.tensor(rng.normal(size=B)) for nm in names}
    psi0 = torch.tensor(rng.normal(size=(B, 2**n)) + 1j * rng.normal(size=(B, 2**n)))
    want = oracle.run(prog, values, psi0).numpy()
    slot_of = {nm: i for i, nm in enumerate(prog.param_names)}
    plan = build_plan(prog.ops, n, slot_of, cfg)
    params = np.stack([values[nm].numpy() for nm in prog.param_names]) if prog.param_names else np.zeros((1, B))
    got = emulate(plan, psi0.numpy(), params)
    assert sorted(set(plan.execution_order())) == list(range(len(prog.ops)))
    np.testing.assert_allclose(got, want, atol=1e-10 * max(1.0, np.abs(want).max()))


def test_hea_fuses_into_few_sweeps() -> None:
    n, depth = 20, 8
    prog = hea_program(n, depth)
    slot_of = {nm: i for i, nm in enumerate(prog.param_names)}
    cfg = TileConfig.default(True, False)
    plan = build_plan(prog.ops, n, slot_of, cfg)
    assert len(prog.ops) == 632
    assert plan.n_sweeps <= 9, plan.n_sweeps  # vs 632 state round trips gate by gate (2^10-amplitude tiles; 8 with 2^11)
    assert build_plan(prog.ops, n, slot_of, TileConfig(m=11, r=4, low=2, fold=3)).n_sweeps <= 8
    assert build_plan(prog.ops, n, slot_of, TileConfig.default(True, True)).n_sweeps <= 9  # adjoint geometry (m = 10, r = 4; first fit: 13)
    assert build_plan(prog.ops, n, slot_of, TileConfig(m=10, r=3, low=2, fold=3, adjoint=True)).n_sweeps <= 8
    assert build_plan(prog.ops, n, slot_of, TileConfig(m=11, r=4, low=2, fold=3), packer="greedy").n_sweeps == 11
    assert int(plan.rounds["n_exec"].sum()) == 0  # no interpreter entries: merged 2x2 + folded CNOT permutations only
    for sw in plan.sweeps:
        assert set(range(cfg.low)) <= set(int(x) for x in sw["tile_bits"][: sw["m"]])  # coalesced low bits


def test_dagger_plan_inverts() -> None:
    rng = np.random.default_rng(7)
    n = 6
    b = ProgramBuilder(n)
    for q in range(n):
        b.RX(q, "a").RY(q, 0.3).RZ(q, "b").H(q)
    for q in range(n - 1):
        b.CNOT(q, q + 1).CPHASE(q, q + 1, "c")
    b.SWAP(0, 5).scale(1.5)
    prog = b.build()
    slot_of = {nm: i for i, nm in enumerate(prog.param_names)}
    plan = build_plan(prog.ops, n, slot_of, TileConfig(m=5, r=3, low=2, fold=3))
    params = rng.normal(size=(3, 2))
    psi0 = rng.normal(size=(2, 2**n)) + 1j * rng.normal(size=(2, 2**n))
    fwd = emulate(plan, psi0, params)
    back = emulate(plan, fwd, params, dagger=True)
    np.testing.assert_allclose(back, psi0 * 1.5**2, atol=1e-12)  # Scale is not unitary: s·conj(s)


@pytest.mark.parametrize("cfg", [TileConfig(m=10, r=3, low=4, fold=3), TileConfig(m=4, r=2, low=2, fold=3)], ids=lambda c: f"m{c.m}r{c.r}")
def test_chain_moment_gradients_match_oracle(cfg: TileConfig) -> None:
    """The adjoint algebra of sweep.cuh (merged chains + four Bloch moments) == the gate-by-gate adjoint."""
    from qadence_b200.program import PauliSum, PauliTerm

    rng = np.random.default_rng(3)
    n, B = 5, 2
    b = ProgramBuilder(n)
    for q in range(n):
        b.RX(q, f"a{q}").RY(q, "shared").RX(q, f"b{q}").PHASE(q, f"p{q}").RZ(q, "shared").H(q).RY(q, f"c{q}")
    for q in range(n - 1):
        b.CNOT(q, q + 1).CRX(q + 1, q, f"cr{q}").CPHASE(q, (q + 2) % n, f"cp{q}").CRZ(q, q + 1, "shared")
    b.scale("s")
    for q in range(n):
        b.RZ(q, f"z{q}").PHASE(q, f"z{q}").RX(q, 0.3).RY(q, f"y{q}")
    prog = b.build()
    obs = PauliSum(n, [PauliTerm(1.0, (), ((q, "Z"),)) for q in range(n)] + [PauliTerm(0.5, (), ((0, "X"), (3, "Y")))])
    values = {nm: torch.tensor(rng.uniform(0.2, 1.5, size=B)) for nm in prog.param_names}
    want_e, want_g = oracle.adjoint_expectation_and_grads(prog, obs, values)
    slot_of = {nm: i for i, nm in enumerate(prog.param_names)}
    plan = build_plan(prog.ops, n, slot_of, cfg)
    assert any(int(g["chain"]) > 1 for g in plan.gates)
    params = np.stack([values[nm].numpy() for nm in prog.param_names])
    psi0 = oracle.zero_state(n, B).numpy()
    psi = emulate(plan, psi0, params)
    lam = oracle.from_pyq_shape(oracle.paulisum_apply(obs, oracle.to_pyq_shape(torch.tensor(psi), n), values)).numpy()
    back_psi, _, grads = emulate_adjoint(plan, psi, lam, params)
    np.testing.assert_allclose(back_psi, psi0, atol=1e-12)
    for nm, i in slot_of.items():
        np.testing.assert_allclose(grads[i], want_g[nm].numpy(), atol=1e-11, err_msg=nm)


@pytest.mark.parametrize("packer", ["search", "greedy"])
@pytest.mark.parametrize("seed", range(4))
def test_tile_search_keeps_the_program_semantics(packer: str, seed: int) -> None:
    """Tiles chosen by the beam search (n > m: several sweeps, windows and postponed frontier gates) execute the same
    program: forward state and adjoint gradients equal the gate-by-gate oracle."""
    from qadence_b200.program import PauliSum, PauliTerm

    rng = np.random.default_rng(40 + seed)
    n, B = 9, 2
    cfg = TileConfig(m=5 + seed % 2, r=2 + seed % 2, low=2, fold=3)
    if seed < 2:
        prog, names = random_program(n, 60, rng)
    else:  # layered circuit: the case the search is built for
        prog = hea_program(n, 3)
        names = prog.param_names
    values = {nm: torch.tensor(rng.uniform(0.2, 1.5, size=B)) for nm in names}
    slot_of = {nm: i for i, nm in enumerate(prog.param_names)}
    plan = build_plan(prog.ops, n, slot_of, cfg, packer=packer)
    assert sorted(set(plan.execution_order())) == list(range(len(prog.ops)))
    params = np.stack([values[nm].numpy() for nm in prog.param_names]) if prog.param_names else np.zeros((1, B))
    psi0 = rng.normal(size=(B, 2**n)) + 1j * rng.normal(size=(B, 2**n))
    want = oracle.run(prog, values, torch.tensor(psi0)).numpy()
    got = emulate(plan, psi0, params)
    np.testing.assert_allclose(got, want, atol=1e-10 * max(1.0, np.abs(want).max()))
    if seed >= 2:
        obs = PauliSum(n, [PauliTerm(1.0, (), ((q, "Z"),)) for q in range(n)])
        want_e, want_g = oracle.adjoint_expectation_and_grads(prog, obs, values)
        z0 = oracle.zero_state(n, B).numpy()
        psi = emulate(plan, z0, params)
        lam = oracle.from_pyq_shape(oracle.paulisum_apply(obs, oracle.to_pyq_shape(torch.tensor(psi), n), values)).numpy()
        back_psi, _, grads = emulate_adjoint(plan, psi, lam, params)
        np.testing.assert_allclose(back_psi, z0, atol=1e-12)
        for nm, i in slot_of.items():
            np.testing.assert_allclose(grads[i], want_g[nm].numpy(), atol=1e-11, err_msg=nm)


def test_tile_search_never_needs_more_sweeps_than_first_fit() -> None:
    rng = np.random.default_rng(11)
    for trial in range(12):
        n = int(rng.integers(7, 13))
        prog, _ = random_program(n, int(rng.integers(30, 120)), rng)
        slot_of = {nm: i for i, nm in enumerate(prog.param_names)}
        cfg = TileConfig(m=int(rng.integers(4, 7)), r=int(rng.integers(2, 4)), low=2, fold=3)

        def cost(plan) -> float:
            return 3.0 * plan.n_sweeps + float(plan.sweeps["n_rounds"].sum())

        a = build_plan(prog.ops, n, slot_of, cfg, packer="greedy")
        b = build_plan(prog.ops, n, slot_of, cfg, packer="search")
        assert cost(b) <= cost(a) + 1e-9, (trial, cost(a), cost(b))


def test_adjoint_walk_is_linear_in_an_arbitrary_cotangent_state() -> None:
    """`engine._RunState` (differentiable `run`) feeds torch's cotangent g of the output state to the adjoint walk as λ and
    halves the result: dL/dθ = Re⟨g|∂ψ/∂θ⟩ = ½·(2 Re⟨λ|∂U|ψ⟩)|λ=g.  Checked here on the emulated plan against torch autograd
    through the oracle for a loss that is NOT an expectation value (λ is not of the form Oψ)."""
    rng = np.random.default_rng(21)
    n, B = 6, 2
    b = ProgramBuilder(n)
    for q in range(n):
        b.RX(q, f"a{q}").RY(q, "shared").PHASE(q, f"p{q}").H(q)
    for q in range(n - 1):
        b.CNOT(q, q + 1).CRZ(q, q + 1, f"c{q}").CPHASE(q + 1, q, "shared")
    b.scale("s")
    for q in range(n):
        b.RZ(q, f"z{q}").RX(q, 0.3)
    prog = b.build()
    values = {nm: torch.tensor(rng.uniform(0.2, 1.5, size=B), requires_grad=True) for nm in prog.param_names}
    psi0 = torch.tensor(rng.normal(size=(B, 2**n)) + 1j * rng.normal(size=(B, 2**n)))
    target = torch.tensor(rng.normal(size=(B, 2**n)) + 1j * rng.normal(size=(B, 2**n)))
    psi = oracle.run(prog, values, psi0)
    loss = ((psi - target).abs() ** 2).sum() + (psi.real * target.imag).sum()  # any real function of the wavefunction
    psi.retain_grad()
    loss.backward()
    g = psi.grad.detach().numpy()  # torch's cotangent: dL/dRe + i dL/dIm
    slot_of = {nm: i for i, nm in enumerate(prog.param_names)}
    plan = build_plan(prog.ops, n, slot_of, TileConfig(m=5, r=3, low=2, fold=3))
    params = np.stack([values[nm].detach().numpy() for nm in prog.param_names])
    back_psi, back_lam, grads = emulate_adjoint(plan, psi.detach().numpy(), g, params)
    np.testing.assert_allclose(back_psi, psi0.numpy(), atol=1e-11)
    for nm, i in slot_of.items():
        np.testing.assert_allclose(0.5 * grads[i], values[nm].grad.numpy(), atol=1e-10, err_msg=nm)
    # U^dagger g is the cotangent of the input state
    psi0r = psi0.clone().requires_grad_(True)
    out = oracle.run(prog, {k: v.detach() for k, v in values.items()}, psi0r)
    (((out - target).abs() ** 2).sum() + (out.real * target.imag).sum()).backward()
    np.testing.assert_allclose(back_lam, psi0r.grad.numpy(), atol=1e-10)
