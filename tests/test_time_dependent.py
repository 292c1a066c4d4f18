"""Time-dependent HamEvo generators (reference: backends/pyqtorch/convert_ops.py:227-231,258-268,
pinned only against qutip to 6e-2 in tests/backends/pyq/test_time_dependent_generator.py:25-62).

The oracle restates the scheme (exponential-midpoint steps with dense exponentials); it is pinned here against an
independent adaptive ODE integration (scipy DOP853) of the reference test's own generator.  The GPU path must match the
oracle to 1e-10 on the same step count, and its continuum adjoint gradients must match finite differences."""

from __future__ import annotations

import numpy as np
import pytest
import scipy.integrate
import torch

from oracle import statevec as oracle
from qadence_b200.program import PauliSum, PauliTerm, Program, ProgramBuilder, total_magnetization

X = np.array([[0, 1], [1, 0]], dtype=complex)
Y = np.array([[0, -1j], [1j, 0]])
Z = np.diag([1.0 + 0j, -1.0])
I2 = np.eye(2, dtype=complex)
OMEGA, FY = 2.0, 3.5  # tests/conftest.py:236-252 of the reference


def reference_test_program(steps: int) -> Program:
    """omega * (sin(y t) X(0) + x t^2 Y(1)) evolved for `duration` (reference tests/conftest.py:261-266)."""
    gen = PauliSum(2, [PauliTerm(OMEGA, (), ((0, "X"),), f"sin({FY}*t)|t"), PauliTerm(OMEGA, (), ((1, "Y"),), "t**2*x|t,x")])
    return ProgramBuilder(2).hamevo_td(gen, "duration", "t", steps).build()


def ode_solution(h, dim: int, T: float, psi0=None) -> np.ndarray:
    y0 = np.zeros(dim, dtype=complex)
    y0[0] = 1.0
    if psi0 is not None:
        y0 = np.asarray(psi0, dtype=complex)
    sol = scipy.integrate.solve_ivp(lambda t, y: -1j * (h(t) @ y), [0.0, T], y0, method="DOP853", rtol=1e-12, atol=1e-13)
    return sol.y[:, -1]


@pytest.mark.parametrize("duration", [0.5, 1.0])
def test_oracle_matches_adaptive_ode_on_the_reference_generator(duration: float) -> None:
    x = 2.0
    prog = Program.loads(reference_test_program(500).dumps())  # through the fixture serialisation too
    out = oracle.run(prog, {"x": torch.tensor([x]), "duration": torch.tensor([duration])})[0].numpy()
    want = ode_solution(lambda t: OMEGA * (np.sin(FY * t) * np.kron(X, I2) + x * t**2 * np.kron(I2, Y)), 4, duration)
    assert np.abs(out - want).max() < 5e-6  # second-order scheme, 500 steps; the reference's own bar is 6e-2
    # halving the step divides the error by four
    e = [np.abs(oracle.run(reference_test_program(s), {"x": torch.tensor([x]), "duration": torch.tensor([duration])})[0].numpy() - want).max()
         for s in (50, 100)]
    assert 3.5 < e[0] / e[1] < 4.5


def test_oracle_batched_durations_and_symbols() -> None:
    prog = reference_test_program(200)
    xs, Ts = [0.3, 1.1, 2.0], [1.0, 0.5, 0.25]
    out = oracle.run(prog, {"x": torch.tensor(xs), "duration": torch.tensor(Ts)}).numpy()
    for b, (x, T) in enumerate(zip(xs, Ts)):
        want = ode_solution(lambda t: OMEGA * (np.sin(FY * t) * np.kron(X, I2) + x * t**2 * np.kron(I2, Y)), 4, T)
        assert np.abs(out[b] - want).max() < 2e-5


# ------------------------------------------------------------------------------------------------ GPU
def _ising_td(n: int, steps: int) -> Program:
    """H(t) = Σ J Z_i Z_{i+1} + a·cos(w t)·Σ X_i + t·b·Y_0 between two layers of rotations."""
    terms = [PauliTerm(0.7, (), ((i, "Z"), (i + 1, "Z"))) for i in range(n - 1)]
    terms += [PauliTerm(1.0, (), ((i, "X"),), "a*cos(w*t)|a,t,w") for i in range(n)]
    terms += [PauliTerm(1.0, (), ((0, "Y"),), "b*t|b,t")]
    b = ProgramBuilder(n)
    for q in range(n):
        b.RY(q, f"th{q}")
    b.hamevo_td(PauliSum(n, terms), "T", "t", steps)
    for q in range(n - 1):
        b.CNOT(q, q + 1)
    for q in range(n):
        b.RX(q, f"ph{q}")
    return b.build()


def _vals(prog: Program, B: int, seed: int) -> dict[str, torch.Tensor]:
    rng = np.random.default_rng(seed)
    return {nm: torch.tensor(rng.uniform(0.2, 1.4, size=B)) for nm in prog.param_names}


@pytest.mark.gpu
def test_gpu_matches_oracle_and_ode() -> None:
    from qadence_b200 import engine

    prog = reference_test_program(500)
    vals = {"x": torch.tensor([2.0, 0.4]), "duration": torch.tensor([1.0, 0.5])}
    got = engine.run(engine.CompiledProgram(prog), vals, device="cuda:0").cpu()
    assert (got - oracle.run(prog, vals)).abs().max() < 1e-10
    want = ode_solution(lambda t: OMEGA * (np.sin(FY * t) * np.kron(X, I2) + 2.0 * t**2 * np.kron(I2, Y)), 4, 1.0)
    assert np.abs(got[0].numpy() - want).max() < 5e-6

    n, B = 5, 3
    prog = _ising_td(n, 40)
    vals = _vals(prog, B, 0)
    got = engine.run(engine.CompiledProgram(prog), vals, device="cuda:0").cpu()
    assert (got - oracle.run(prog, vals)).abs().max() < 1e-10
    assert (got.abs().pow(2).sum(1) - 1).abs().max() < 1e-12


@pytest.mark.gpu
def test_gpu_diagonal_time_dependent_generator() -> None:
    """All-Z generators take the on-the-fly diagonal kernel; the midpoint rule is then exact for polynomial-in-t
    coefficients of degree ≤ 1 and the result equals exp(−i ∫H dt)."""
    from qadence_b200 import engine

    n = 6
    gen = PauliSum(n, [PauliTerm(0.8, (), ((i, "Z"), ((i + 1) % n, "Z")), "c*t|c,t") for i in range(n)] + [PauliTerm(0.3, (), ((2, "Z"),))])
    b = ProgramBuilder(n)
    for q in range(n):
        b.H(q)
    prog = b.hamevo_td(gen, "T", "t", 7).build()
    vals = {"c": torch.tensor([0.9, 1.7], dtype=torch.float64), "T": torch.tensor([1.3, 0.6], dtype=torch.float64)}
    got = engine.run(engine.CompiledProgram(prog), vals, device="cuda:0").cpu()
    assert (got - oracle.run(prog, vals)).abs().max() < 1e-10
    # closed form: ∫_0^T c t dt = c T²/2
    static = PauliSum(n, [PauliTerm(0.8, ("k",), ((i, "Z"), ((i + 1) % n, "Z"))) for i in range(n)] + [PauliTerm(0.3, ("T",), ((2, "Z"),))])
    b2 = ProgramBuilder(n)
    for q in range(n):
        b2.H(q)
    exact = oracle.run(b2.hamevo(static, 1.0).build(), {"k": vals["c"] * vals["T"] ** 2 / 2, "T": vals["T"]})
    assert (got - exact).abs().max() < 1e-10


@pytest.mark.gpu
def test_gpu_adjoint_gradients_through_time_dependent_evolution() -> None:
    from qadence_b200 import engine

    n, B = 4, 2
    prog = _ising_td(n, 120)
    obs = total_magnetization(n)
    vals = {k: v.requires_grad_(True) for k, v in _vals(prog, B, 1).items()}
    cp = engine.CompiledProgram(prog)
    e = engine.expectation(cp, [engine.CompiledObservable(obs)], vals, device="cuda:0")
    det = {k: v.detach().clone() for k, v in vals.items()}
    assert (e.detach().cpu() - oracle.expectation(prog, [obs], det)).abs().max() < 1e-10
    e.sum().backward()
    h = 1e-5
    for nm in ("th1", "ph2", "T", "a", "w", "b"):
        for bidx in range(B):
            up, dn = {k: v.clone() for k, v in det.items()}, {k: v.clone() for k, v in det.items()}
            up[nm][bidx] += h
            dn[nm][bidx] -= h
            fd = (oracle.expectation(prog, [obs], up)[bidx, 0] - oracle.expectation(prog, [obs], dn)[bidx, 0]) / (2 * h)
            # gate angles: exact adjoint of the discretised circuit; T and the coefficient symbols: continuum formulas, O(dt²)
            tol = 1e-7 if nm in ("th1", "ph2") else 5e-4
            assert abs(vals[nm].grad[bidx].item() - fd.item()) < tol, (nm, bidx, vals[nm].grad[bidx].item(), fd.item())


# --------------------------------------------------------------------------- Trotter option (algo_hevo="trotter")
def _rydberg_like(n: int):
    from qadence_b200.program import PauliSum, PauliTerm

    terms = []
    for i in range(n):
        for j in range(i + 1, n):
            terms.append(PauliTerm(0.9 / (1 + abs(i - j)) ** 3, (), ((i, "Z"), (j, "Z"))))
        terms += [PauliTerm(0.4 + 0.1 * i, (), ((i, "Z"),)), PauliTerm(0.8, ("om",), ((i, "X"),)), PauliTerm(-0.3 - 0.05 * i, (), ((i, "Y"),))]
    return PauliSum(n, terms)


def test_trotter_restatement_converges_to_the_exact_propagator() -> None:
    """CPU: the Strang scheme the oracle restates has an error ∝ 1/k² against the exact exponential."""
    from qadence_b200.program import Op, OpKind, Program

    n, B = 4, 2
    gen = _rydberg_like(n)
    vals = {"t": torch.tensor([0.7, 1.1]), "om": torch.tensor([1.0, 1.3])}
    psi0 = torch.tensor(np.random.default_rng(0).normal(size=(B, 2**n)) + 0j)
    psi0 = psi0 / psi0.norm(dim=1, keepdim=True)
    exact = oracle.run(Program(n, [Op(OpKind.HAMEVO, tuple(range(n)), (), param="t", generator=gen)]), vals, psi0)
    errs = []
    for k in (8, 16, 32):
        got = oracle.run(Program(n, [Op(OpKind.HAMEVO, tuple(range(n)), (), param="t", generator=gen, steps=k)]), vals, psi0)
        errs.append((got - exact).abs().max().item())
    assert errs[0] / errs[1] > 3.5 and errs[1] / errs[2] > 3.5 and errs[2] < 2e-3, errs


@pytest.mark.gpu
@pytest.mark.parametrize("dagger", [False, True])
def test_trotter_on_the_gpu_equals_the_restatement(dagger: bool) -> None:
    """GPU: `hamevo.evolve_trotter` (on-the-fly diagonal kernel + one fused sweep of drive rotations per step) against the
    oracle's dense restatement of the SAME scheme at equal step count, 1e-10; and within O(dt²) of the exact propagator."""
    from qadence_b200 import engine, hamevo
    from qadence_b200.program import Op, OpKind, Program

    n, B, k = 8, 3, 20
    gen = _rydberg_like(n)
    rng = np.random.default_rng(1)
    vals = {"t": torch.tensor(rng.uniform(0.3, 1.2, size=B)), "om": torch.tensor(rng.uniform(0.5, 1.5, size=B))}
    psi0 = torch.tensor(rng.normal(size=(B, 2**n)) + 1j * rng.normal(size=(B, 2**n)))
    psi0 = psi0 / psi0.norm(dim=1, keepdim=True)
    op = Op(OpKind.HAMEVO, tuple(range(n)), (), param="t", generator=gen, steps=k)
    want = oracle.from_pyq_shape(oracle.hamevo_apply(op, oracle.to_pyq_shape(psi0, n), vals, n, dagger=dagger))
    cps = engine.CompiledPauliSum.build(gen)
    got = hamevo.evolve(op, cps, psi0.to("cuda:0").clone(), vals, n, dagger=dagger).cpu()
    assert (got - want).abs().max().item() < 1e-10
    exact_op = Op(OpKind.HAMEVO, tuple(range(n)), (), param="t", generator=gen)
    exact = oracle.from_pyq_shape(oracle.hamevo_apply(exact_op, oracle.to_pyq_shape(psi0, n), vals, n, dagger=dagger))
    assert (got - exact).abs().max().item() < 5e-2
    # through the program path (engine.run picks the Trotter branch from op.steps)
    out = engine.run(engine.CompiledProgram(Program(n, [op])), vals, psi0, device="cuda:0").cpu()
    fwd = oracle.run(Program(n, [op]), vals, psi0)
    assert (out - fwd).abs().max().item() < 1e-10
