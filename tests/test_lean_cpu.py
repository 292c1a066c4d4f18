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


# ------------------------------------------------------------------------------------------ tracked permutations
def test_affine_maps_compose_and_invert() -> None:
    rng = np.random.default_rng(11)
    m = 7
    for _ in range(20):
        # a random invertible GF(2) matrix as a product of transvections (CNOTs) and swaps, plus constants / external vectors
        a = P._Affine.identity(m)
        for _ in range(12):
            i, j = (int(x) for x in rng.choice(m, size=2, replace=False))
            step = P._Affine([(1 << k) | ((1 << j) if k == i else 0) for k in range(m)], int(rng.integers(1 << m)), {int(rng.integers(20, 24)): int(rng.integers(1 << m))})
            a = step.after(a)
        inv = a.inverse()
        for ext_mask in (0, 1, 5):
            def apply(f, x):
                v = f.lin(x) ^ f.const
                for k, (e, vec) in enumerate(sorted(f.ext.items())):
                    if (ext_mask >> (e - 20)) & 1:
                        v ^= vec
                return v
            for x in rng.integers(0, 1 << m, size=16):
                assert apply(inv, apply(a, int(x))) == int(x)
                both = inv.after(a)
                assert apply(both, int(x)) == int(x)


@pytest.mark.parametrize("adjoint", [False, True])
def test_tracked_permutations_equal_physical_ones_and_drop_the_barriers(any_geometry, monkeypatch, adjoint: bool) -> None:
    """`B200SV_VPERM=1` (default) and `=0` execute the same program on the emulator; with tracking, only the first and the last
    round of a TMA-staged sweep keep the barrier between their loads and stores."""
    n, B = 9, 2
    prog = hea_program(n, 3)
    rng = np.random.default_rng(2)
    values = {nm: torch.tensor(rng.uniform(0.1, 3.0, size=B)) for nm in prog.param_names}
    slot_of = {nm: i for i, nm in enumerate(prog.param_names)}
    params = np.stack([values[nm].numpy() for nm in prog.param_names])
    cfg = TileConfig(m=6, r=3, low=2, fold=3, adjoint=adjoint)
    plans = {}
    for v in ("1", "0"):
        monkeypatch.setenv("B200SV_VPERM", v)
        plans[v] = build_plan(prog.ops, n, slot_of, cfg)
    tabs = {v: (p.lean_tab_bwd if adjoint else p.lean_tab_fwd) for v, p in plans.items()}
    assert not np.array_equal(tabs["1"], tabs["0"])
    sync = {v: 0 for v in tabs}
    inner = {v: 0 for v in tabs}
    for v, plan in plans.items():
        for sw in plan.sweeps:
            if int(sw["io_mode"]) == P.IO_GENERIC:
                continue
            fr, nr = int(sw["first_round"]), int(sw["n_rounds"])
            for k in range(nr):
                s_ = int(tabs[v][fr + k][37]) & 1
                sync[v] += s_
                if 0 < k < nr - 1:
                    inner[v] += s_
        assert check_tables_hazard_free(plan, not adjoint) > 0
    assert inner["1"] == 0 and inner["0"] > 0 and sync["1"] < sync["0"]
    z0 = oracle.zero_state(n, B).numpy()
    want = oracle.run(prog, values).numpy()
    if not adjoint:
        for v, plan in plans.items():
            psi, _, _ = emulate_lean(plan, z0, params)
            np.testing.assert_allclose(psi, want, atol=1e-11, err_msg=f"VPERM={v}")
    else:
        obs = PauliSum(n, [PauliTerm(1.0, (), ((q, "Z"),)) for q in range(n)] + [PauliTerm(0.3, (), ((0, "X"), (4, "Y")))])
        _, want_g = oracle.adjoint_expectation_and_grads(prog, obs, values)
        lam = oracle.from_pyq_shape(oracle.paulisum_apply(obs, oracle.to_pyq_shape(torch.tensor(want), n), values)).numpy()
        for v, plan in plans.items():
            back_psi, _, grads = emulate_lean(plan, want, params, lam=lam)
            np.testing.assert_allclose(back_psi, z0, atol=1e-11, err_msg=f"VPERM={v}")
            for nm, i in slot_of.items():
                np.testing.assert_allclose(grads[i], want_g[nm].numpy(), atol=1e-10, err_msg=f"VPERM={v} {nm}")


def test_low_lanes_stay_bank_conflict_free_under_tracked_permutations(any_geometry) -> None:
    """Quarter warps (complex128: 8 lanes × 16 bytes) must hit 8 different 16-byte columns on the load AND the store side of
    every round of cfg2's production tables, although the tracked permutations make the address vectors of the thread bits
    arbitrary: `_lean_table` re-chooses which positions sit on the low lane bits."""
    n = 20
    prog = hea_program(n, 8)
    slot_of = {nm: i for i, nm in enumerate(prog.param_names)}
    for adjoint in (False, True):
        plan = build_plan(prog.ops, n, slot_of, TileConfig(m=10, r=4, low=2, fold=3, adjoint=adjoint))
        tab = plan.lean_tab_bwd if adjoint else plan.lean_tab_fwd
        bad = total = 0
        for sw in plan.sweeps:
            if int(sw["io_mode"]) == P.IO_GENERIC:
                continue
            for k in range(int(sw["n_rounds"])):
                T = tab[int(sw["first_round"]) + k]
                lanes = np.arange(8)
                w = T[lanes & 15]
                for half, shift in ((0, 0), (1, 16)):
                    cols = ((w >> shift) & 0xFFFF) >> 4 & 7
                    total += 1
                    bad += len(set(int(c) for c in cols)) != 8
        assert bad <= total // 8, (adjoint, bad, total)  # 8 of 96 today: rounds that face the un-swizzled TMA layout with low register bits


def test_register_bit_fallbacks_for_sweeps_the_lean_kernel_cannot_take() -> None:
    """A plan that needs the generic kernel (controlled rotations, or a register smaller than the tile) never carries r = 5
    (the generic kernel holds 4 register bits), and complex128 adjoint plans fall back to the r = 3 geometry it was tuned for."""
    from qadence_b200 import cabi

    try:
        cabi.load()
    except cabi.B200svError:
        pytest.skip("libb200sv.so not built")
    n = 14
    b = ProgramBuilder(n)
    for q in range(n):
        b.RX(q, f"a{q}")
    for q in range(n - 1):
        b.CRZ(q, q + 1, f"c{q}")
    prog = b.build()
    slot_of = {nm: i for i, nm in enumerate(prog.param_names)}
    p64 = build_plan(prog.ops, n, slot_of, TileConfig.default(False, False))
    assert TileConfig.default(False, False).r == 5 and all(int(s["r"]) <= 4 for s in p64.sweeps)
    pad = build_plan(prog.ops, n, slot_of, TileConfig.default(True, True))
    assert TileConfig.default(True, True).r == 4 and all(int(s["r"]) == 3 for s in pad.sweeps)
    small = hea_program(5, 1)
    ps = build_plan(small.ops, 5, {nm: i for i, nm in enumerate(small.param_names)}, TileConfig.default(False, False))
    assert all(int(s["r"]) <= 4 and int(s["m"]) - int(s["r"]) >= 0 for s in ps.sweeps)
    # a pure HEA keeps the lean geometry
    hp = hea_program(n, 2)
    ph = build_plan(hp.ops, n, {nm: i for i, nm in enumerate(hp.param_names)}, TileConfig.default(True, True))
    assert all(int(s["r"]) == 4 for s in ph.sweeps)
