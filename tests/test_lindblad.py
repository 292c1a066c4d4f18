"""Lindblad `noise_operators` of HamEvo (backends/pyqtorch/convert_ops.py:249-267): the master equation
dρ/dt = −i[H, ρ] + Σ_k (L_k ρ L_k† − ½{L_k†L_k, ρ}).

Three independent statements of the same map are compared:
  * the ODE itself, integrated on the density MATRIX by scipy's DOP853 to 1e-12 (no vectorisation identity involved);
  * the oracle's restatement (`oracle.statevec.lindblad_apply`: exp(t𝓛) on the whole doubled register);
  * the product path (`hamevo.evolve_lindblad`: exp(t𝓛) on the 2k support qubits, `b200sv_apply_dense`) — GPU tests.
The reference itself pins this case only against qutip `mesolve` to 6e-2 (tests/backends/pyq/test_time_dependent_generator.py:68-126).
"""
from __future__ import annotations

import numpy as np
import pytest
import torch
from scipy.integrate import solve_ivp

from oracle import statevec as oracle
from qadence_b200 import density, hamevo
from qadence_b200.program import Op, OpKind, PauliSum, PauliTerm, Program, ProgramBuilder

X = np.array([[0, 1], [1, 0]], dtype=np.complex128)
Y = np.array([[0, -1j], [1j, 0]], dtype=np.complex128)
Z = np.array([[1, 0], [0, -1]], dtype=np.complex128)
SM = np.array([[0, 1], [0, 0]], dtype=np.complex128)  # |0><1|: decay of the excited state |1>
PAULI = {"X": X, "Y": Y, "Z": Z}


def kron_all(ms):
    out = np.ones((1, 1), dtype=np.complex128)
    for m in ms:
        out = np.kron(out, m)
    return out


def embed_on(m: np.ndarray, support, n: int) -> np.ndarray:
    """Operator `m` on the qubits `support` (support[0] = MSB of m) of an n-qubit register, qubit 0 = MSB."""
    k = len(support)
    rest = [q for q in range(n) if q not in support]
    full = np.kron(m, np.eye(2 ** (n - k)))  # on the order support + rest
    order = list(support) + rest
    t = full.reshape([2] * (2 * n))
    perm = [order.index(q) for q in range(n)]
    t = t.transpose(perm + [n + p for p in perm])
    return t.reshape(2**n, 2**n)


def dense_h(ps: PauliSum, coef) -> np.ndarray:
    n = ps.n_qubits
    H = np.zeros((2**n, 2**n), dtype=np.complex128)
    for c, term in zip(coef, ps.terms):
        ms = [np.eye(2, dtype=np.complex128)] * n
        for q, p in term.paulis:
            ms[q] = PAULI[p]
        H += c * kron_all(ms)
    return H


def integrate_master_equation(h_of_t, jumps, rho0: np.ndarray, T: float) -> np.ndarray:
    d = rho0.shape[0]

    def rhs(t, y):
        rho = y.reshape(d, d)
        H = h_of_t(t)
        out = -1j * (H @ rho - rho @ H)
        for L in jumps:
            LL = L.conj().T @ L
            out += L @ rho @ L.conj().T - 0.5 * (LL @ rho + rho @ LL)
        return out.reshape(-1)

    sol = solve_ivp(rhs, (0.0, T), rho0.astype(np.complex128).reshape(-1), method="DOP853", rtol=1e-12, atol=1e-14)
    return sol.y[:, -1].reshape(d, d)


def run_vectorised(prog: Program, values: dict, state=None) -> torch.Tensor:
    vprog, needs = density.vectorise(prog)
    n = prog.n_qubits
    v = oracle.run(vprog, density.extend_values(values, needs), density.vectorise_state(state, n))
    return v.reshape(v.shape[0], 2**n, 2**n)


def random_state(rng, n, B=1):
    s = rng.normal(size=(B, 2**n)) + 1j * rng.normal(size=(B, 2**n))
    return torch.as_tensor(s / np.linalg.norm(s, axis=1, keepdims=True))


def lindblad_case(seed: int):
    """3 qubits, a Rydberg-like generator on qubits (0, 2) with a batched coefficient, two jump operators, gates around it."""
    rng = np.random.default_rng(seed)
    n = 3
    gen = PauliSum(n, [PauliTerm(0.8, ("w",), ((0, "X"),)), PauliTerm(-0.45, (), ((2, "Y"),)), PauliTerm(1.3, (), ((0, "Z"), (2, "Z"))),
                       PauliTerm(0.3, (), ((2, "Z"),))])
    L1 = np.sqrt(0.4) * np.kron(SM, np.eye(2))  # decay of qubit 0 (first of the support (0, 2))
    L2 = 0.5 * rng.normal(size=(4, 4)) + 0.5j * rng.normal(size=(4, 4))  # a generic two-qubit jump operator
    b = ProgramBuilder(n)
    b.RX(0, "a").H(1).CNOT(1, 2).RY(2, 0.7)
    pre = b.build().ops
    evo = Op(OpKind.HAMEVO, (0, 2), (), param="t", generator=gen, label="HamEvo", jumps=(L1, L2))
    post = ProgramBuilder(n).RZ(1, "a").CNOT(0, 1).build().ops
    prog = Program(n, [*pre, evo, *post])
    values = {"a": torch.tensor([0.3, -1.1]), "w": torch.tensor([1.0, 0.6]), "t": torch.tensor([0.9, 1.7])}
    return prog, values, gen, (L1, L2), pre, post


def direct_density(prog_case, b: int) -> np.ndarray:
    prog, values, gen, jumps, pre, post = prog_case
    n = prog.n_qubits
    vb = {k: v[b:b + 1] for k, v in values.items()}
    psi = oracle.run(Program(n, pre), vb)[0].numpy()
    rho = np.outer(psi, psi.conj())
    coef = [0.8 * float(values["w"][b]), -0.45, 1.3, 0.3]
    H = dense_h(gen, coef)
    Ls = [embed_on(L, (0, 2), n) for L in jumps]
    rho = integrate_master_equation(lambda t: H, Ls, rho, float(values["t"][b]))
    U = oracle.run(Program(n, post), vb, torch.eye(2**n, dtype=torch.complex128)).numpy().T
    return U @ rho @ U.conj().T


def test_amplitude_damping_has_the_textbook_solution() -> None:
    """H = 0, L = √γ σ⁻: ρ11(t) = ρ11(0) e^{−γt}, coherences decay with e^{−γt/2}."""
    gamma, t = 0.7, 1.3
    evo = Op(OpKind.HAMEVO, (0,), (), param=t, generator=PauliSum(1, [PauliTerm(0.0, (), ((0, "Z"),))]), jumps=(np.sqrt(gamma) * SM,))
    psi0 = torch.tensor([[0.6, 0.8j]], dtype=torch.complex128)
    rho = run_vectorised(Program(1, [evo]), {}, psi0)[0].numpy()
    assert abs(rho[1, 1] - 0.64 * np.exp(-gamma * t)) < 1e-13
    assert abs(rho[0, 1] - (0.6 * np.conj(0.8j)) * np.exp(-gamma * t / 2)) < 1e-13
    assert abs(np.trace(rho) - 1) < 1e-13


def test_dissipator_of_the_product_equals_the_definition() -> None:
    rng = np.random.default_rng(3)
    Ls = [rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4)) for _ in range(3)]
    rho = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
    want = sum(L @ rho @ L.conj().T - 0.5 * (L.conj().T @ L @ rho + rho @ L.conj().T @ L) for L in Ls)
    got = (hamevo.lindblad_dissipator(Ls) @ rho.reshape(-1)).reshape(4, 4)
    assert np.abs(got - want).max() < 1e-13


@pytest.mark.parametrize("seed", range(3))
def test_oracle_restatement_equals_the_integrated_master_equation(seed: int) -> None:
    case = lindblad_case(seed)
    prog, values = case[0], case[1]
    assert prog.is_noisy
    got = run_vectorised(prog, values).numpy()
    for b in range(2):
        want = direct_density(case, b)
        assert np.abs(got[b] - want).max() < 5e-10  # the ODE solver's tolerance, not the restatement's
        assert abs(np.trace(got[b]) - 1) < 1e-12
        assert np.linalg.eigvalsh((got[b] + got[b].conj().T) / 2).min() > -1e-12


def td_case():
    n = 2
    gen = PauliSum(n, [PauliTerm(1.0, (), ((0, "X"),), "sin(0.9*t)*x|t,x"), PauliTerm(0.5, (), ((1, "Y"),), "cos(t)*t|t"),
                       PauliTerm(0.7, (), ((0, "Z"), (1, "Z")))])
    L = 0.6 * kron_all([X, np.eye(2)])
    evo = Op(OpKind.HAMEVO, (0, 1), (), param="duration", generator=gen, label="HamEvo(t)", tsym="t", steps=400, jumps=(L,))
    values = {"x": torch.tensor([1.3]), "duration": torch.tensor([1.1])}
    return Program(n, [evo]), values, gen, L


def test_time_dependent_lindblad_restatement_converges_to_the_ode() -> None:
    prog, values, gen, L = td_case()
    got = run_vectorised(prog, values)[0].numpy()

    def h(t):
        return dense_h(gen, [np.sin(0.9 * t) * 1.3, 0.5 * np.cos(t) * t, 0.7])

    rho0 = np.zeros((4, 4), dtype=np.complex128)
    rho0[0, 0] = 1
    want = integrate_master_equation(h, [L], rho0, 1.1)
    assert np.abs(got - want).max() < 1e-5  # exponential midpoint rule: O(dt²) with dt = 1.1 / 400


def test_lindblad_ops_round_trip_through_json() -> None:
    prog = lindblad_case(0)[0]
    again = Program.from_json(prog.to_json())
    op = [o for o in again.ops if o.kind == OpKind.HAMEVO][0]
    assert len(op.jumps) == 2 and np.allclose(op.jumps[1], lindblad_case(0)[3][1])
    assert again.is_noisy


@pytest.mark.parametrize("case", ["constant", "time-dependent"])
def test_host_side_of_the_product_propagator(monkeypatch, case: str) -> None:
    """`hamevo.evolve_lindblad` (support-local superoperator, term order of the COMPILED Pauli sum, batch handling) with the
    one device call it makes, `engine.apply_dense`, replaced by the oracle's dense operator application."""
    from qadence_b200 import engine as E
    from qadence_b200.program import OpKind as K

    def apply_dense(psi, matrix, bits):
        n = int(psi.shape[1]).bit_length() - 1
        m = torch.as_tensor(matrix).to(torch.complex128)
        out = oracle.apply_operator(oracle.to_pyq_shape(psi, n), m[:, :, None], [n - 1 - b for b in bits])
        return oracle.from_pyq_shape(out).contiguous()

    monkeypatch.setattr(E, "apply_dense", apply_dense)
    prog, values = (lindblad_case(1) if case == "constant" else td_case())[:2]
    n = prog.n_qubits
    vprog, needs = density.vectorise(prog)
    vals = density.extend_values(values, needs)
    B = max(int(torch.as_tensor(v).numel()) for v in values.values())
    state = density.vectorise_state(oracle.zero_state(n, B), n)
    for op in vprog.ops:
        if op.kind == K.HAMEVO and op.jumps:
            state = hamevo.evolve_lindblad(op, E.CompiledPauliSum.build(op.generator), state, vals, 2 * n)
        else:
            state = oracle.run(Program(2 * n, [op]), vals, state)
    want = run_vectorised(prog, values).reshape(B, -1)
    assert (state - want).abs().max() < 1e-12


# ------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(3))
def test_gpu_lindblad_equals_the_restatement(seed: int) -> None:
    prog, values = lindblad_case(seed)[:2]
    want = run_vectorised(prog, values)
    cd = density.CompiledDensityProgram(prog)
    assert not cd.compiled.invertible
    got = density.run(cd, values, device="cuda").cpu()
    assert (got - want).abs().max() < 1e-12
    # a mixed initial state
    rho0 = 0.5 * density.vectorise_state(random_state(np.random.default_rng(seed), 3), 3).reshape(1, 8, 8)
    rho0 = rho0 + 0.5 * torch.eye(8, dtype=torch.complex128)[None] / 8
    want = run_vectorised(prog, values, rho0)
    got = density.run(cd, values, rho0.cuda(), device="cuda").cpu()
    assert (got - want).abs().max() < 1e-12


@pytest.mark.gpu
def test_gpu_time_dependent_lindblad_equals_the_restatement() -> None:
    prog, values = td_case()[:2]
    want = run_vectorised(prog, values)
    got = density.run(density.CompiledDensityProgram(prog), values, device="cuda").cpu()
    assert (got - want).abs().max() < 1e-11


@pytest.mark.gpu
def test_gpu_lindblad_expectation_is_the_trace() -> None:
    from qadence_b200 import engine

    prog, values = lindblad_case(1)[:2]
    obs = PauliSum(3, [PauliTerm(1.0, (), ((q, "Z"),)) for q in range(3)] + [PauliTerm(0.5, (), ((0, "X"), (2, "X")))])
    rho = run_vectorised(prog, values).numpy()
    O = dense_h(obs, [1.0, 1.0, 1.0, 0.5])
    want = np.array([np.trace(O @ rho[b]).real for b in range(2)])
    cd = density.CompiledDensityProgram(prog)
    got = density.expectation(cd, [engine.CompiledObservable(density.lift_observable(obs, 3))], values, device="cuda").cpu().numpy()[:, 0]
    assert np.abs(got - want).max() < 1e-12
