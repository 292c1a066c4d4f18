"""GPU parity tests (run with `-m gpu` on a B200): the CUDA path, called through the C-ABI, against
the CPU oracle on identical seeded inputs.  Tolerances are BASELINE.json's: 1e-10 for complex128,
1e-5 for complex64, bit-exact for sampled indices given identical uniforms."""

from __future__ import annotations

import numpy as np
import pytest
import torch

from oracle import statevec as oracle
from qadence_b200 import engine
from qadence_b200.plan import TileConfig
from qadence_b200.program import (
    OpKind,
    PauliSum,
    PauliTerm,
    ProgramBuilder,
    hea_program,
    total_magnetization,
)
from tests.test_plan_cpu import random_program

pytestmark = pytest.mark.gpu

TOL = {torch.complex128: 1e-10, torch.complex64: 1e-5}


def _rand_state(rng, B, n):
    s = rng.normal(size=(B, 2**n)) + 1j * rng.normal(size=(B, 2**n))
    s /= np.linalg.norm(s, axis=1, keepdims=True)
    return torch.tensor(s)


@pytest.mark.parametrize("dtype", [torch.complex128, torch.complex64])
@pytest.mark.parametrize("n,batch", [(1, 1), (2, 3), (3, 2), (5, 4), (8, 3), (11, 2), (12, 1), (13, 2), (15, 1)])
def test_random_circuit_run(n: int, batch: int, dtype: torch.dtype) -> None:
    rng = np.random.default_rng(1000 + 10 * n + batch)
    prog, names = random_program(n, 60, rng)
    # keep the state norm O(1): the generator inserts non-unitary scales / random matrices
    values = {nm: torch.tensor(rng.uniform(0.5, 1.5, size=batch)) for nm in names}
    psi0 = _rand_state(rng, batch, n)
    want = oracle.run(prog, values, psi0)
    cp = engine.CompiledProgram(prog)
    got = engine.run(cp, values, psi0, dtype=dtype, device="cuda:0").cpu().to(torch.complex128)
    scale = max(1.0, want.abs().max().item())
    assert (got - want).abs().max().item() <= TOL[dtype] * scale


@pytest.mark.parametrize("cfg", [TileConfig(11, 3, 4, 3), TileConfig(10, 3, 4, 3), TileConfig(12, 4, 5, 3), TileConfig(9, 2, 3, 3), TileConfig(8, 1, 2, 3),
                                 TileConfig(2, 1, 0, 4), TileConfig(3, 2, 0, 3)], ids=lambda c: f"m{c.m}r{c.r}")  # the last two: fewer threads than rounds
def test_tile_geometries(cfg: TileConfig) -> None:
    """Every (m, r) instantiation of the sweep kernel, through the C-ABI directly."""
    import ctypes as C

    from qadence_b200 import cabi
    from qadence_b200.engine import DevicePlan, _DT, _stream_ptr
    from qadence_b200.plan import build_plan

    n, B = 13, 2
    rng = np.random.default_rng(cfg.m * 10 + cfg.r)
    prog, names = random_program(n, 50, rng)
    values = {nm: torch.tensor(rng.uniform(0.5, 1.5, size=B)) for nm in names}
    psi0 = _rand_state(rng, B, n)
    want = oracle.run(prog, values, psi0)
    dev = torch.device("cuda:0")
    slot_of = {nm: i for i, nm in enumerate(prog.param_names)}
    dp = DevicePlan(build_plan(prog.ops, n, slot_of, cfg), dev)
    tab = engine.pack_params(prog.param_names, values, B, dev)
    st = psi0.to(dev).clone()
    rc = cabi.load().b200sv_apply_plan(st.data_ptr(), _DT[st.dtype], n, B, tab.data_ptr(), C.byref(dp.struct), 0, dp.n_sweeps, 0, dp.workspace(st.dtype, B, False).data_ptr(), _stream_ptr(dev))
    assert rc == 0
    assert (st.cpu() - want).abs().max().item() <= 1e-10 * max(1.0, want.abs().max().item())


@pytest.mark.parametrize("dtype", [torch.complex128, torch.complex64])
@pytest.mark.parametrize("n,depth,batch", [(4, 2, 16), (6, 2, 4), (11, 2, 3), (13, 1, 2)])
def test_hea_expectation_and_adjoint(n: int, depth: int, batch: int, dtype: torch.dtype) -> None:
    prog = hea_program(n, depth)
    obs = PauliSum(n, [PauliTerm(1.0, (), ((q, "Z"),)) for q in range(n)] + [PauliTerm(0.7, (), ((0, "X"), (n - 1, "Y"))), PauliTerm(-0.3, (), ((0, "Y"),))])
    g = torch.Generator().manual_seed(n * 100 + depth)
    values = {nm: (2 * torch.pi * torch.rand(batch, generator=g, dtype=torch.float64)).requires_grad_(True) for nm in prog.param_names}
    cp = engine.CompiledProgram(prog)
    out = engine.expectation(cp, [engine.CompiledObservable(obs), engine.CompiledObservable(total_magnetization(n))], values, dtype=dtype, device="cuda:0")
    w = torch.tensor([1.0, -0.5], device=out.device, dtype=out.dtype)
    (out * w).sum().backward()
    det = {k: v.detach() for k, v in values.items()}
    e1, g1 = oracle.adjoint_expectation_and_grads(prog, obs, det)
    e2, g2 = oracle.adjoint_expectation_and_grads(prog, total_magnetization(n), det)
    tol = TOL[dtype] * (1 if dtype == torch.complex128 else 10)
    assert (out[:, 0].detach().cpu().double() - e1).abs().max().item() <= tol
    assert (out[:, 1].detach().cpu().double() - e2).abs().max().item() <= tol
    for k in values:
        want = g1[k] - 0.5 * g2[k]
        assert (values[k].grad.cpu() - want).abs().max().item() <= tol * 10, k


def test_second_backward_over_a_retained_graph() -> None:
    """The first backward pass consumes the saved wavefunction (the adjoint walk un-applies the circuit in place); a second
    pass over a retained graph recomputes it and must give the same gradients again."""
    n, batch = 9, 3
    prog = hea_program(n, 2)
    g = torch.Generator().manual_seed(5)
    values = {nm: (2 * torch.pi * torch.rand(batch, generator=g, dtype=torch.float64)).requires_grad_(True) for nm in prog.param_names}
    out = engine.expectation(engine.CompiledProgram(prog), [engine.CompiledObservable(total_magnetization(n))], values, device="cuda:0")
    out.sum().backward(retain_graph=True)
    first = {k: v.grad.clone() for k, v in values.items()}
    out.sum().backward()
    _, want = oracle.adjoint_expectation_and_grads(prog, total_magnetization(n), {k: v.detach() for k, v in values.items()})
    for k in values:
        assert (first[k].cpu() - want[k]).abs().max().item() <= 1e-10, k
        assert (values[k].grad.cpu() - 2 * want[k]).abs().max().item() <= 2e-10, k


def test_shared_and_scalar_parameters() -> None:
    """`[1]`-shaped (shared) parameters: autograd sum-reduces the per-batch gradients."""
    n, batch = 5, 6
    b = ProgramBuilder(n)
    for q in range(n):
        b.RX(q, "x").RY(q, "theta").CPHASE(q, (q + 1) % n, "phi").CRZ(q, (q + 2) % n, "theta")
    b.scale("s")
    prog = b.build()
    obs = total_magnetization(n)
    x = torch.linspace(0.1, 1.2, batch, dtype=torch.float64).requires_grad_(True)
    theta = torch.tensor([0.37], dtype=torch.float64, requires_grad=True)
    phi = torch.tensor([1.1], dtype=torch.float64, requires_grad=True)
    s = torch.tensor([1.3], dtype=torch.float64, requires_grad=True)
    values = {"x": x, "theta": theta, "phi": phi, "s": s}
    cp = engine.CompiledProgram(prog)
    out = engine.expectation(cp, [engine.CompiledObservable(obs)], values, device="cuda:0")
    out.sum().backward()
    e, g = oracle.adjoint_expectation_and_grads(prog, obs, {k: v.detach() for k, v in values.items()})
    assert (out[:, 0].detach().cpu() - e).abs().max().item() < 1e-10
    assert (x.grad - g["x"]).abs().max().item() < 1e-9
    for nm, t in (("theta", theta), ("phi", phi), ("s", s)):
        assert abs(t.grad.item() - g[nm].sum().item()) < 1e-9, nm


def test_pauli_apply_and_dense_and_permute() -> None:
    n, B = 9, 3
    rng = np.random.default_rng(5)
    psi = _rand_state(rng, B, n)
    terms = [PauliTerm(0.3 - 0.2j, (), ((0, "X"), (3, "Y"), (8, "Z"))), PauliTerm(1.1, ("a",), ((2, "Z"), (4, "Z"))), PauliTerm(-0.4, (), ()), PauliTerm(0.9, ("a", "b"), ((8, "Y"),))]
    ps = PauliSum(n, terms)
    values = {"a": torch.tensor(rng.normal(size=B)), "b": torch.tensor([0.7])}
    want = oracle.from_pyq_shape(oracle.paulisum_apply(ps, oracle.to_pyq_shape(psi, n), values))
    cps = engine.CompiledPauliSum.build(ps)
    dpsi = psi.to("cuda:0")
    got = engine.pauli_apply(cps, cps.coef(values, dpsi.device), dpsi).cpu()
    assert (got - want).abs().max().item() < 1e-12
    e = engine.pauli_expectation(cps, cps.coef(values, dpsi.device), dpsi).cpu()
    assert (e - torch.einsum("bi,bi->b", psi.conj(), want).real).abs().max().item() < 1e-12
    # dense 3-qubit matrix on scattered targets
    m = rng.normal(size=(8, 8)) + 1j * rng.normal(size=(8, 8))
    targets = [6, 1, 4]
    want_d = oracle.from_pyq_shape(oracle.apply_operator(oracle.to_pyq_shape(psi, n), torch.tensor(m)[:, :, None], targets))
    got_d = engine.apply_dense(dpsi, m, [n - 1 - q for q in targets]).cpu()
    assert (got_d - want_d).abs().max().item() < 1e-12
    # endianness inversion == qadence.transpile.invert_endianness(Tensor)
    assert torch.equal(engine.bit_reverse(dpsi).cpu(), oracle.invert_endianness_state(psi, n))


@pytest.mark.parametrize("n", [1, 4, 8, 10, 13])
def test_sampling_bit_exact(n: int) -> None:
    B, shots = 3, 500
    rng = np.random.default_rng(n)
    psi = _rand_state(rng, B, n)
    u = torch.tensor(rng.random(size=(B, shots)))
    idx = engine.sample_indices(psi.to("cuda:0"), u).cpu().numpy()
    want = oracle.sample_indices(oracle.probabilities(psi), u.numpy())
    assert np.array_equal(idx, want)
    psi32 = psi.to(torch.complex64)
    idx32 = engine.sample_indices(psi32.to("cuda:0"), u).cpu().numpy()
    assert np.array_equal(idx32, oracle.sample_indices(oracle.probabilities(psi32), u.numpy()))


def test_sample_counters_match_reference_kats() -> None:
    """`tests/backends/pyq/test_quantum_pyq.py:620-632`: X→{'1':10}, Y→{'1':10}, Z→{'0':10}."""
    for gate, key in (("X", "1"), ("Y", "1"), ("Z", "0")):
        cp = engine.CompiledProgram(ProgramBuilder(1).gate(gate, 0).build())
        assert engine.sample(cp, n_shots=10, device="cuda:0") == [{key: 10}]


def test_hamevo_krylov_and_diagonal() -> None:
    n, B = 6, 4
    rng = np.random.default_rng(11)
    zz = [PauliTerm(float(rng.normal()), (), ((i, "Z"), (j, "Z"))) for i in range(n) for j in range(i + 1, n)]
    drive = [PauliTerm(0.8, (), ((q, "X"),)) for q in range(n)] + [PauliTerm(-0.3, (), ((q, "Y"),)) for q in range(n)]
    diag = PauliSum(n, zz + [PauliTerm(0.5, (), ((q, "Z"),)) for q in range(n)])
    full = PauliSum(n, zz + drive)
    psi = _rand_state(rng, B, n)
    values = {"t": torch.tensor(rng.uniform(0.1, 3.0, size=B))}
    for gen in (diag, full):
        prog = ProgramBuilder(n).H(0).hamevo(gen, "t").CNOT(0, 1).build()
        want = oracle.run(prog, values, psi)
        got = engine.run(engine.CompiledProgram(prog), values, psi, device="cuda:0").cpu()
        assert (got - want).abs().max().item() < 1e-10
    # dense generator on a qubit subset + adjoint gradient w.r.t. the evolution time
    g = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
    g = g + g.conj().T
    prog = ProgramBuilder(n).RX(2, "a").hamevo(g, "t", targets=(4, 1)).hamevo(full, "t").build()
    vals = {"t": values["t"].clone().requires_grad_(True), "a": torch.tensor([0.4], requires_grad=True, dtype=torch.float64)}
    obs = total_magnetization(n)
    out = engine.expectation(engine.CompiledProgram(prog), [engine.CompiledObservable(obs)], vals, device="cuda:0")
    out.sum().backward()
    e, gr = oracle.adjoint_expectation_and_grads(prog, obs, {k: v.detach() for k, v in vals.items()})
    assert (out[:, 0].detach().cpu() - e).abs().max().item() < 1e-10
    assert (vals["t"].grad - gr["t"]).abs().max().item() < 1e-8
    assert abs(vals["a"].grad.item() - gr["a"].sum().item()) < 1e-8


def test_sharded_simulator_world_1_matches_engine() -> None:
    """The qubit-sharded driver with a single rank (g = 0) is the plain engine."""
    import numpy as np

    from qadence_b200 import dist as qdist
    from qadence_b200.program import Program
    from tests.test_dist_cpu import qft_program

    n = 12
    prog = Program(n, qft_program(n).ops + hea_program(n, 1).ops)
    g = torch.Generator().manual_seed(3)
    values = {nm: 2 * torch.pi * torch.rand(1, generator=g, dtype=torch.float64) for nm in prog.param_names}
    sim = qdist.QubitShardedSimulator(prog, 0, 1, "cuda:0")
    st = sim.run(values)
    want = oracle.run(prog, values)
    assert (st.cpu() - want).abs().max().item() < 1e-10
    assert abs(sim.norm2(st).item() - 1.0) < 1e-10
    e = sim.expectation_diagonal(st, total_magnetization(n)).item()
    assert abs(e - oracle.expectation(prog, [total_magnetization(n)], values).item()) < 1e-10


def test_randomised_differential_fuzz() -> None:
    """A short run of tools/gpu_fuzz.py (random circuits x random tile geometries, forward and adjoint) — the campaign that
    found the tiny-CTA bug (per-tile permutation constants of rounds beyond the CTA's thread count)."""
    import subprocess
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parents[1]
    out = subprocess.run([sys.executable, str(root / "tools" / "gpu_fuzz.py"), "12", "5"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0 and out.stdout.strip().startswith("ok:"), out.stdout[-1500:] + out.stderr[-1500:]


def test_randomised_differential_fuzz_of_the_other_kernels() -> None:
    """A short run of tools/gpu_fuzz_ops.py: Pauli sums, diagonal / Krylov HamEvo, dense matrices, bit-exact sampling."""
    import subprocess
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parents[1]
    out = subprocess.run([sys.executable, str(root / "tools" / "gpu_fuzz_ops.py"), "10", "3"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0 and out.stdout.strip().startswith("ok:"), out.stdout[-1500:] + out.stderr[-1500:]


@pytest.mark.parametrize("tool,seconds", [("gpu_fuzz_mixed.py", "10"), ("gpu_fuzz_embedding.py", "6")])
def test_randomised_differential_fuzz_of_whole_programs_and_embedding(tool: str, seconds: str) -> None:
    import subprocess
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parents[1]
    out = subprocess.run([sys.executable, str(root / "tools" / tool), seconds, "4"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0 and out.stdout.strip().startswith("ok:"), out.stdout[-1500:] + out.stderr[-1500:]
