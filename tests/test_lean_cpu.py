"""Lean sweeps on the CPU: the host-built index tables, TMA boxes, pivot flips and the normalised-2x2 algebra of
`lean_sweep_kernel`, emulated step by step (tests/lean_emulator.py), must reproduce the gate-by-gate oracle."""

from __future__ import annotations

import numpy as np
import pytest
import torch

from oracle import statevec as oracle
from qadence_b200 import plan as P
from qadence_b200.plan import TileConfig, build_plan
from qadence_b200.program import CONST_MATS, OpKind, PauliSum, PauliTerm, ProgramBuilder, hea_program
from tests.lean_emulator import check_tables_hazard_free, emulate_lean


@pytest.fixture()
def any_geometry(monkeypatch):
    """Pretend the lean kernel exists for every (m, r): the tables are geometry-generic, only the set of compiled
    instantiations is not."""
    monkeypatch.setattr(P, "_lean_supported", lambda c128, m, r, adjoint: True)


def lean_random_program(n: int, n_gates: int, rng: np.random.Generator, n_params: int = 6, unitary: bool = False):
    """Only what a lean sweep can hold: uncontrolled 1-qubit gates and X / CNOT / SWAP permutations."""
    b = ProgramBuilder(n)
    names = [f"p{i}" for i in range(n_params)]
    fixed = [k for k in CONST_MATS if k not in ("Zero",)]
    if unitary:
        fixed = [k for k in fixed if k in ("X", "Y", "Z", "H", "S", "T", "SDagger")]
    for _ in range(n_gates):
        kind = int(rng.integers(0, 8))
        qs = [int(q) for q in rng.permutation(n)]
        t = qs[0]
        param = names[int(rng.integers(n_params))] if rng.random() < 0.7 else float(rng.normal())
        if kind == 0:
            b.gate(fixed[int(rng.integers(len(fixed)))], t, [])
        elif kind in (1, 2, 3, 4):
            b.rot([OpKind.RX, OpKind.RY, OpKind.RZ, OpKind.PHASE][kind - 1], t, param, [])
        elif kind == 5 and n >= 2:
            b.CNOT(qs[0], qs[1])
        elif kind == 6 and n >= 2:
            b.SWAP(qs[0], qs[1])
        elif not unitary:
            m = rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2))
            if rng.random() < 0.3:
                m[int(rng.integers(2))] = 0  # singular: a zero row (projector-like)
            b._add(__import__("qadence_b200").Op(OpKind.MAT1, (t,), (), matrix=m, label="rand"))
        else:
            b.X(t)
    return b.build(), names


GEOMS = [(5, 2, 2, 3), (6, 3, 2, 3), (7, 4, 2, 3), (6, 2, 1, 4), (8, 3, 2, 4), (7, 5, 2, 3), (6, 3, 0, 3)]


@pytest.mark.parametrize("geom", GEOMS, ids=lambda g: f"m{g[0]}r{g[1]}low{g[2]}f{g[3]}")
@pytest.mark.parametrize("n", [8, 9])
def test_lean_forward_matches_oracle(any_geometry, geom, n: int) -> None:
    m, r, low, fold = geom
    rng = np.random.default_rng(1000 * n + 10 * m + r)
    prog, names = lean_random_program(n, 70, rng)
    B = 2
    values = {nm: torch.tensor(rng.normal(size=B)) for nm in names}
    psi0 = rng.normal(size=(B, 2**n)) + 1j * rng.normal(size=(B, 2**n))
    want = oracle.run(prog, values, torch.tensor(psi0)).numpy()
    slot_of = {nm: i for i, nm in enumerate(prog.param_names)}
    plan = build_plan(prog.ops, n, slot_of, TileConfig(m=m, r=r, low=low, fold=fold))
    assert all(int(s["io_mode"]) != P.IO_GENERIC for s in plan.sweeps)
    params = np.stack([values[nm].numpy() for nm in prog.param_names]) if prog.param_names else np.zeros((1, B))
    got, _, _ = emulate_lean(plan, psi0, params)
    np.testing.assert_allclose(got, want, atol=1e-10 * max(1.0, np.abs(want).max()))


@pytest.mark.parametrize("geom", GEOMS[:5], ids=lambda g: f"m{g[0]}r{g[1]}low{g[2]}f{g[3]}")
@pytest.mark.parametrize("layered", [False, True])
def test_lean_adjoint_matches_oracle(any_geometry, geom, layered: bool) -> None:
    m, r, low, fold = geom
    n, B = 8, 2
    rng = np.random.default_rng(77 + m + 10 * r + int(layered))
    if layered:
        prog = hea_program(n, 3)
        names = prog.param_names
    else:
        prog, names = lean_random_program(n, 60, rng, n_params=8, unitary=True)
    values = {nm: torch.tensor(rng.uniform(0.1, 3.0, size=B)) for nm in names}
    obs = PauliSum(n, [PauliTerm(1.0, (), ((q, "Z"),)) for q in range(n)] + [PauliTerm(0.5, (), ((0, "X"), (3, "Y")))])
    want_e, want_g = oracle.adjoint_expectation_and_grads(prog, obs, values)
    slot_of = {nm: i for i, nm in enumerate(prog.param_names)}
    fplan = build_plan(prog.ops, n, slot_of, TileConfig(m=m, r=r, low=low, fold=fold))
    aplan = build_plan(prog.ops, n, slot_of, TileConfig(m=max(m - 1, r + 1), r=r, low=low, fold=fold, adjoint=True))
    params = np.stack([values[nm].numpy() for nm in prog.param_names])
    z0 = oracle.zero_state(n, B).numpy()
    psi, _, _ = emulate_lean(fplan, z0, params)
    np.testing.assert_allclose(psi, oracle.run(prog, values).numpy(), atol=1e-11)
    lam = oracle.from_pyq_shape(oracle.paulisum_apply(obs, oracle.to_pyq_shape(torch.tensor(psi), n), values)).numpy()
    back_psi, back_lam, grads = emulate_lean(aplan, psi, params, lam=lam)
    np.testing.assert_allclose(back_psi, z0, atol=1e-11)
    for nm, i in slot_of.items():
        if nm in want_g:
            np.testing.assert_allclose(grads[i], want_g[nm].numpy(), atol=1e-10, err_msg=nm)


@pytest.mark.parametrize("c128", [True, False])
def test_lean_with_hardware_swizzled_tma_layout(any_geometry, monkeypatch, c128: bool) -> None:
    """B200SV_TMA_SWIZZLE=1: the first / last round of a sweep address the tile in the layout CU_TENSOR_MAP_SWIZZLE_128B
    produces; forward state and adjoint gradients still equal the oracle."""
    monkeypatch.setenv("B200SV_TMA_SWIZZLE", "1")
    n, B = 10, 2
    rng = np.random.default_rng(9)
    prog = hea_program(n, 3)
    values = {nm: torch.tensor(rng.uniform(0.1, 3.0, size=B)) for nm in prog.param_names}
    obs = PauliSum(n, [PauliTerm(1.0, (), ((q, "Z"),)) for q in range(n)])
    want_e, want_g = oracle.adjoint_expectation_and_grads(prog, obs, values)
    slot_of = {nm: i for i, nm in enumerate(prog.param_names)}
    fold = 3 if c128 else 4
    fplan = build_plan(prog.ops, n, slot_of, TileConfig(m=8, r=3, low=2, fold=fold))
    aplan = build_plan(prog.ops, n, slot_of, TileConfig(m=7, r=3, low=2, fold=fold, adjoint=True))
    assert any(int(t["rank"]) & P.TMA_SWIZZLE_FLAG for t in fplan.tma) and any(int(t["rank"]) & P.TMA_SWIZZLE_FLAG for t in aplan.tma)
    params = np.stack([values[nm].numpy() for nm in prog.param_names])
    z0 = oracle.zero_state(n, B).numpy()
    psi, _, _ = emulate_lean(fplan, z0, params)
    np.testing.assert_allclose(psi, oracle.run(prog, values).numpy(), atol=1e-11)
    lam = oracle.from_pyq_shape(oracle.paulisum_apply(obs, oracle.to_pyq_shape(torch.tensor(psi), n), values)).numpy()
    back_psi, _, grads = emulate_lean(aplan, psi, params, lam=lam)
    np.testing.assert_allclose(back_psi, z0, atol=1e-11)
    for nm, i in slot_of.items():
        np.testing.assert_allclose(grads[i], want_g[nm].numpy(), atol=1e-10, err_msg=nm)


def test_production_geometry_marks_hea_sweeps_lean_and_tma() -> None:
    """cfg2's plans with the library's real instantiation list: every sweep lean, the windowed tiles TMA boxes."""
    from qadence_b200 import cabi

    try:
        cabi.load()
    except cabi.B200svError:
        pytest.skip("libb200sv.so not built")
    n = 20
    prog = hea_program(n, 8)
    slot_of = {nm: i for i, nm in enumerate(prog.param_names)}
    for adjoint in (False, True):
        plan = build_plan(prog.ops, n, slot_of, TileConfig.default(True, adjoint))
        modes = [int(s["io_mode"]) for s in plan.sweeps]
        assert all(md != P.IO_GENERIC for md in modes), modes
        assert sum(md == P.IO_LEAN_TMA for md in modes) >= plan.n_sweeps - 2, modes
        for s, tm in zip(plan.sweeps, plan.tma):
            if int(s["io_mode"]) == P.IO_LEAN_TMA:
                rank = int(tm["rank"]) & 0xFF
                assert 2 <= rank <= 5
                for d in range(rank):
                    bits = int(tm["nbits"][d]) + (1 if d == 0 else 0)
                    assert not tm["is_tile"][d] or bits <= 8  # box dimensions hold at most 256 8-byte elements


def test_tma_box_enumerates_the_tile() -> None:
    assert P.tma_box([0, 1, 5, 6, 7], 10, True) is not None
    assert P.tma_box([1, 2, 3], 10, True) is None  # bit 0 outside the tile: innermost row of 16 bytes only
    assert P.tma_box([0, 2, 4, 6, 8], 10, True) is None  # 5 runs: rank 10
    box = P.tma_box(list(range(11)), 20, True)[0]  # contiguous tile: split at 256 elements
    assert int(box["rank"]) == 3 and list(box["nbits"][:3]) == [7, 4, 9] and list(box["is_tile"][:3]) == [1, 1, 0]
    box = P.tma_box([0, 1] + list(range(11, 20)), 20, True)[0]
    assert int(box["rank"]) == 5 and list(box["is_tile"][:5]) == [1, 0, 1, 1, 0] and int(box["nbits"][4]) == 0


def test_small_production_plan_on_the_emulator() -> None:
    """hea(12, 2) with the production complex128 geometries (the instantiations the library really has)."""
    from qadence_b200 import cabi

    try:
        cabi.load()
    except cabi.B200svError:
        pytest.skip("libb200sv.so not built")
    n, B = 12, 1
    prog = hea_program(n, 2)
    rng = np.random.default_rng(5)
    values = {nm: torch.tensor(rng.uniform(0.1, 3.0, size=B)) for nm in prog.param_names}
    slot_of = {nm: i for i, nm in enumerate(prog.param_names)}
    params = np.stack([values[nm].numpy() for nm in prog.param_names])
    obs = PauliSum(n, [PauliTerm(1.0, (), ((q, "Z"),)) for q in range(n)])
    want_e, want_g = oracle.adjoint_expectation_and_grads(prog, obs, values)
    fplan = build_plan(prog.ops, n, slot_of, TileConfig.default(True, False))
    aplan = build_plan(prog.ops, n, slot_of, TileConfig.default(True, True))
    assert all(int(s["io_mode"]) != P.IO_GENERIC for s in list(fplan.sweeps) + list(aplan.sweeps))
    z0 = oracle.zero_state(n, B).numpy()
    psi, _, _ = emulate_lean(fplan, z0, params)
    np.testing.assert_allclose(psi, oracle.run(prog, values).numpy(), atol=1e-11)
    lam = oracle.from_pyq_shape(oracle.paulisum_apply(obs, oracle.to_pyq_shape(torch.tensor(psi), n), values)).numpy()
    back_psi, _, grads = emulate_lean(aplan, psi, params, lam=lam)
    np.testing.assert_allclose(back_psi, z0, atol=1e-11)
    for nm, i in slot_of.items():
        np.testing.assert_allclose(grads[i], want_g[nm].numpy(), atol=1e-10, err_msg=nm)


@pytest.mark.parametrize("shape", ["cfg2", "cfg4"])
def test_production_tables_are_hazard_free(shape: str) -> None:
    """Static shared-memory race check of the tables the GPU runs for configs[1] / configs[3] (both directions, both dtypes):
    see `check_tables_hazard_free` — compute-sanitizer's racecheck is closed on the GPU pool (profiles/r02_hygiene.md)."""
    from qadence_b200 import cabi
    from qadence_b200.program import Program

    try:
        cabi.load()
    except cabi.B200svError:
        pytest.skip("libb200sv.so not built")
    if shape == "cfg2":
        n, prog = 20, hea_program(20, 8)
    else:
        n = 16
        b = ProgramBuilder(n)
        for q in range(n):
            b.RX(q, "x")
        prog = Program(n, b.build().ops + hea_program(n, 12).ops)
    slot_of = {nm: i for i, nm in enumerate(prog.param_names)}
    total = 0
    for c128 in (True, False):
        for adjoint in (False, True):
            plan = build_plan(prog.ops, n, slot_of, TileConfig.default(c128, adjoint))
            assert all(int(s["io_mode"]) != P.IO_GENERIC for s in plan.sweeps)
            total += check_tables_hazard_free(plan, forward=not adjoint)
    assert total > 100
