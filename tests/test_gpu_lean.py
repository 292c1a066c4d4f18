"""`lean_sweep_kernel` (TMA-staged tiles, persistent CTAs, normalised 2x2) against the CPU oracle, through the C-ABI:
every compiled (dtype, r, m, direction) instantiation, both tile I/O modes (TMA boxes and cp.async gathers), non-unitary / singular 1-qubit matrices (row flips), and the adjoint gradients."""

from __future__ import annotations

import ctypes as C

import numpy as np
import pytest
import torch

from oracle import statevec as oracle
from qadence_b200 import cabi, engine
from qadence_b200 import plan as P
from qadence_b200.engine import DevicePlan, _DT, _stream_ptr
from qadence_b200.plan import TileConfig, build_plan
from qadence_b200.program import PauliSum, PauliTerm, hea_program
from tests.test_lean_cpu import lean_random_program

pytestmark = pytest.mark.gpu

TOL = {torch.complex128: 1e-10, torch.complex64: 1e-5}
# (c128, r, m, adjoint): mirrors B200SV_LEAN_INSTANCES in csrc/b200sv.cu (test_instance_list_matches_library checks it)
INSTANCES = [
    (True, 4, 11, False), (True, 5, 12, False), (True, 4, 10, False), (True, 4, 10, True), (True, 3, 10, True), (True, 4, 11, True), (True, 3, 11, True),
    (False, 4, 12, False), (False, 5, 12, False), (False, 4, 11, False), (False, 3, 11, True), (False, 4, 12, True), (False, 4, 11, True),
]


def _rand_state(rng, B, n):
    s = rng.normal(size=(B, 2**n)) + 1j * rng.normal(size=(B, 2**n))
    s /= np.linalg.norm(s, axis=1, keepdims=True)
    return torch.tensor(s)


def test_instance_list_matches_library() -> None:
    lib = cabi.load()
    for c128 in (True, False):
        for adj in (False, True):
            for m in range(4, 14):
                for r in range(1, 6):
                    want = (c128, r, m, adj) in INSTANCES
                    assert bool(lib.b200sv_lean_supported(cabi.C128 if c128 else cabi.C64, m, r, int(adj))) == want, (c128, r, m, adj)


def _forward(prog, values, psi0, cfg, dtype, n, B):
    dev = torch.device("cuda:0")
    slot_of = {nm: i for i, nm in enumerate(prog.param_names)}
    plan = build_plan(prog.ops, n, slot_of, cfg)
    dp = DevicePlan(plan, dev)
    tab = engine.pack_params(prog.param_names, values, B, dev)
    st = psi0.to(device=dev, dtype=dtype).clone()
    rc = cabi.load().b200sv_apply_plan(st.data_ptr(), _DT[dtype], n, B, tab.data_ptr(), C.byref(dp.struct), 0, dp.n_sweeps, 0,
                                       dp.workspace(dtype, B, False).data_ptr(), _stream_ptr(dev))
    assert rc == 0
    torch.cuda.synchronize()
    return st.cpu().to(torch.complex128), plan


@pytest.mark.parametrize("inst", [i for i in INSTANCES if not i[3]], ids=lambda i: f"{'c128' if i[0] else 'c64'}-r{i[1]}m{i[2]}")
@pytest.mark.parametrize("io", ["tma", "gather"])
def test_lean_forward_instances(inst, io: str, monkeypatch) -> None:
    c128, r, m, _ = inst
    dtype = torch.complex128 if c128 else torch.complex64
    monkeypatch.setenv("B200SV_TMA", "0" if io == "gather" else "1")
    n, B = m + 1, 3  # one outer bit: every tile has at most two runs, i.e. is a TMA box
    rng = np.random.default_rng(31 * m + r + len(io))
    prog, names = lean_random_program(n, 90, rng, unitary=not c128)  # complex64: keep the norm O(1)
    values = {nm: torch.tensor(rng.uniform(0.2, 2.5, size=B)) for nm in names}
    psi0 = _rand_state(rng, B, n)
    want = oracle.run(prog, values, psi0)
    got, plan = _forward(prog, values, psi0, TileConfig(m=m, r=r, low=2, fold=3 if c128 else 4), dtype, n, B)
    modes = [int(s["io_mode"]) for s in plan.sweeps]
    assert all(md != P.IO_GENERIC for md in modes), modes
    boxes = [P.tma_box(sw["tile_bits"][:m], n, c128) is not None for sw in plan.sweeps]
    assert modes == [(P.IO_LEAN_TMA if (bx and io != "gather") else P.IO_LEAN) for bx in boxes], (modes, boxes)
    scale = max(1.0, want.abs().max().item())
    assert (got - want).abs().max().item() <= TOL[dtype] * scale


@pytest.mark.parametrize("inst", [i for i in INSTANCES if i[3]], ids=lambda i: f"{'c128' if i[0] else 'c64'}-r{i[1]}m{i[2]}")
@pytest.mark.parametrize("io", ["tma", "gather"])
def test_lean_adjoint_instances(inst, io: str, monkeypatch) -> None:
    """ψ, λ walked back and all gradients, through `b200sv_adjoint_plan` directly with this instantiation's geometry."""
    c128, r, m, _ = inst
    dtype = torch.complex128 if c128 else torch.complex64
    monkeypatch.setenv("B200SV_TMA", "0" if io == "gather" else "1")
    n, B = m + 1, 2
    rng = np.random.default_rng(17 * m + r)
    prog = hea_program(n, 2)
    values = {nm: torch.tensor(rng.uniform(0.1, 3.0, size=B)) for nm in prog.param_names}
    obs = PauliSum(n, [PauliTerm(1.0, (), ((q, "Z"),)) for q in range(n)] + [PauliTerm(0.5, (), ((0, "X"), (3, "Y")))])
    want_e, want_g = oracle.adjoint_expectation_and_grads(prog, obs, values)
    psi = oracle.run(prog, values)
    lam = oracle.from_pyq_shape(oracle.paulisum_apply(obs, oracle.to_pyq_shape(psi, n), values))
    dev = torch.device("cuda:0")
    slot_of = {nm: i for i, nm in enumerate(prog.param_names)}
    plan = build_plan(prog.ops, n, slot_of, TileConfig(m=m, r=r, low=2, fold=3 if c128 else 4, adjoint=True))
    assert all(int(s["io_mode"]) != P.IO_GENERIC for s in plan.sweeps)
    dp = DevicePlan(plan, dev)
    tab = engine.pack_params(prog.param_names, values, B, dev)
    dpsi, dlam = psi.to(device=dev, dtype=dtype).clone(), lam.to(device=dev, dtype=dtype).clone()
    grads = torch.zeros(len(prog.param_names), B, dtype=torch.float64, device=dev)
    rc = cabi.load().b200sv_adjoint_plan(dpsi.data_ptr(), dlam.data_ptr(), _DT[dtype], n, B, tab.data_ptr(), grads.data_ptr(),
                                         C.byref(dp.struct), 0, dp.n_sweeps, 0, dp.workspace(dtype, B, True).data_ptr(), _stream_ptr(dev))
    assert rc == 0
    torch.cuda.synchronize()
    tol = TOL[dtype]
    z0 = oracle.zero_state(n, B)
    assert (dpsi.cpu().to(torch.complex128) - z0).abs().max().item() <= tol
    for nm, i in slot_of.items():
        assert (grads[i].cpu() - want_g[nm]).abs().max().item() <= 10 * tol, nm


@pytest.mark.parametrize("dtype", [torch.complex128, torch.complex64])
@pytest.mark.parametrize("n,depth,batch", [(12, 3, 5), (14, 2, 3)])
def test_lean_engine_expectation_and_gradients(n: int, depth: int, batch: int, dtype: torch.dtype) -> None:
    """The default geometries through the engine (what `Backend.expectation(...).backward()` runs)."""
    prog = hea_program(n, depth)
    obs = PauliSum(n, [PauliTerm(1.0, (), ((q, "Z"),)) for q in range(n)] + [PauliTerm(0.7, (), ((0, "X"), (n - 1, "Y")))])
    g = torch.Generator().manual_seed(n * 100 + depth)
    values = {nm: (2 * torch.pi * torch.rand(batch, generator=g, dtype=torch.float64)).requires_grad_(True) for nm in prog.param_names}
    cp = engine.CompiledProgram(prog)
    seg = cp.segments[0]
    for adj in (False, True):
        plan = cp.device_plan(seg, dtype, adj, torch.device("cuda:0")).plan
        assert all(int(s["io_mode"]) != P.IO_GENERIC for s in plan.sweeps)
    out = engine.expectation(cp, [engine.CompiledObservable(obs)], values, dtype=dtype, device="cuda:0")
    out.sum().backward()
    want_e, want_g = oracle.adjoint_expectation_and_grads(prog, obs, {k: v.detach() for k, v in values.items()})
    tol = TOL[dtype]
    assert (out[:, 0].detach().cpu() - want_e).abs().max().item() <= 10 * tol
    for k in values:
        assert (values[k].grad.cpu() - want_g[k]).abs().max().item() <= 20 * tol, k


@pytest.mark.parametrize("batch", [1, 7])
def test_lean_odd_batches_and_partial_ranges(batch: int) -> None:
    """Non-power-of-two batch (TMA top dimension = batch x high bits) and per-sweep ranges (each sweep carries its own
    pending scalar, so any [first, last) split gives the same state)."""
    n = 13
    rng = np.random.default_rng(batch)
    prog = hea_program(n, 3)
    values = {nm: torch.tensor(rng.uniform(0.1, 3.0, size=batch)) for nm in prog.param_names}
    want = oracle.run(prog, values)
    dev = torch.device("cuda:0")
    slot_of = {nm: i for i, nm in enumerate(prog.param_names)}
    dp = DevicePlan(build_plan(prog.ops, n, slot_of, TileConfig.default(True, False)), dev)
    tab = engine.pack_params(prog.param_names, values, batch, dev)
    st = engine.zero_state(n, batch, torch.complex128, dev)
    ws = dp.workspace(torch.complex128, batch, False)
    for s in range(dp.n_sweeps):
        rc = cabi.load().b200sv_apply_plan(st.data_ptr(), _DT[st.dtype], n, batch, tab.data_ptr(), C.byref(dp.struct), s, s + 1, 0, ws.data_ptr(), _stream_ptr(dev))
        assert rc == 0
    assert (st.cpu() - want).abs().max().item() <= 1e-10
