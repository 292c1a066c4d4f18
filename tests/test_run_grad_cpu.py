"""Plumbing of the differentiable `engine.run` (`_RunState`) on CPU: the three C-ABI-backed helpers it composes are replaced
by oracle doubles with the SAME contracts (`_adjoint_walk` returns 2 Re<λ|dU|ψ>, ψ_in, U†λ), so argument order, the ½ factor,
shared / batched parameters, input-state gradients and dtype / device handling are exercised.  The walk itself is checked
against the oracle by tests/test_plan_cpu.py (emulated plan) and on the GPU by tests/test_gpu_parity.py."""

from __future__ import annotations

import numpy as np
import pytest
import torch

from oracle import statevec as oracle
from qadence_b200 import engine
from qadence_b200.program import ProgramBuilder


@pytest.fixture
def doubles(monkeypatch: pytest.MonkeyPatch) -> None:
    cpu = torch.device("cpu")
    monkeypatch.setattr(engine, "_dev", lambda device: cpu)
    monkeypatch.setattr(engine, "prepare_state", lambda n, batch, state, dtype, device: (
        oracle.zero_state(n, batch).to(dtype) if state is None else state.detach().to(dtype).expand(batch, -1).clone()))

    def values_of(cp, ptable):
        return {nm: ptable[i] for i, nm in enumerate(cp.names)}

    def run_segments(cp, st, ptable, values, index_hi=0):
        return oracle.run(cp.program, values_of(cp, ptable.detach()), st.to(torch.complex128)).to(st.dtype)

    def adjoint_walk(cp, values, B, ptable, psi, lam, grad_names=None):
        # contract: g[slot, b] = 2 Re<λ_b| dU/dθ[slot, b] |ψ_in,b>,  ψ_in,  U†λ
        with torch.enable_grad():
            vals = {nm: ptable[i].detach().clone().requires_grad_(True) for i, nm in enumerate(cp.names)}
        inv = oracle.run_dagger(cp.program, {k: v.detach() for k, v in vals.items()}, psi.to(torch.complex128)) if hasattr(oracle, "run_dagger") else None
        if inv is None:  # invert by solving: the test programs are unitary up to Scale, U†U = |s|² 1
            raise RuntimeError
        psi_in = inv
        with torch.enable_grad():  # (backward passes run with grad mode off)
            st = psi_in.detach().clone().requires_grad_(True)
            out = oracle.run(cp.program, vals, st)
            f = 2.0 * (lam.to(torch.complex128).conj() * out).sum().real
            gs = torch.autograd.grad(f, [*vals.values(), st], allow_unused=True)
        grads = torch.stack([torch.zeros(B, dtype=torch.float64) if g_ is None else g_.expand(B).clone() for g_ in gs[:-1]]) if vals else torch.zeros(1, B)
        return grads, psi_in, 0.5 * gs[-1]

    monkeypatch.setattr(engine, "_run_segments", run_segments)
    monkeypatch.setattr(engine, "_adjoint_walk", adjoint_walk)


def _program():
    b = ProgramBuilder(3)
    b.RX(0, "a").RY(1, "b").CNOT(0, 1).RZ(2, "a").CRX(1, 2, "c").H(0).PHASE(2, 0.4)
    return b.build()


def test_run_is_differentiable_wrt_parameters_and_state(doubles: None, monkeypatch: pytest.MonkeyPatch) -> None:
    prog = _program()

    def run_dagger(program, values, psi):  # U is unitary here: U† = U⁻¹, obtained from the dense matrix
        n = program.n_qubits
        outs = []
        for b in range(psi.shape[0]):
            vb = {k: (v[b:b + 1] if v.numel() > 1 else v) for k, v in values.items()}
            u = oracle.run(program, vb, torch.eye(2**n, dtype=torch.complex128)).T
            outs.append(u.conj().T @ psi[b])
        return torch.stack(outs)

    monkeypatch.setattr(oracle, "run_dagger", run_dagger, raising=False)
    cp = engine.CompiledProgram(prog)
    rng = np.random.default_rng(0)
    B = 3
    vals = {"a": torch.tensor(rng.uniform(0, 3, size=B), requires_grad=True), "b": torch.tensor([0.7], requires_grad=True),
            "c": torch.tensor(rng.uniform(0, 3, size=B))}
    psi0 = torch.tensor(rng.normal(size=(1, 8)) + 1j * rng.normal(size=(1, 8)))
    psi0 = (psi0 / psi0.norm()).requires_grad_(True)
    target = torch.tensor(rng.normal(size=(B, 8)) + 1j * rng.normal(size=(B, 8)))

    def loss_of(psi):
        return ((psi - target).abs() ** 2).sum() + (psi.imag * target.real).sum()

    out = engine.run(cp, vals, psi0)
    assert out.requires_grad and out.shape == (B, 8)
    loss_of(out).backward()
    got = {k: v.grad.clone() for k, v in vals.items() if v.requires_grad}
    got_state = psi0.grad.clone()
    ref_vals = {k: v.detach().clone().requires_grad_(v.requires_grad) for k, v in vals.items()}
    ref_state = psi0.detach().clone().requires_grad_(True)
    loss_of(oracle.run(prog, ref_vals, ref_state.expand(B, -1))).backward()
    for k in got:
        assert torch.allclose(got[k], ref_vals[k].grad, atol=1e-10), k
    assert got_state.shape == (1, 8) and torch.allclose(got_state, ref_state.grad, atol=1e-10)
    # no gradient requested -> plain tensor, same numbers
    with torch.no_grad():
        plain = engine.run(cp, vals, psi0)
    assert not plain.requires_grad and torch.allclose(plain, out.detach())
    # complex64 state in, gradient comes back in the state's dtype
    s32 = psi0.detach().to(torch.complex64).requires_grad_(True)
    o32 = engine.run(cp, {k: v.detach() for k, v in vals.items()}, s32, dtype=torch.complex64)
    o32.abs().pow(2).sum().backward()
    assert s32.grad.dtype == torch.complex64 and s32.grad.shape == (1, 8)


def test_adjoint_walk_through_non_unitary_matrix_ops(monkeypatch: pytest.MonkeyPatch) -> None:
    """Noise superoperators (density-matrix path) are dense, invertible, NON-unitary ops: the walk must un-apply ψ with M⁻¹
    and move λ with M†; singular ops (projectors) must be refused.  `apply_dense` is doubled by the oracle."""
    from qadence_b200 import density
    from qadence_b200.program import Op, OpKind, Program

    def apply_dense(psi, matrix, target_bits):
        n = int(psi.shape[1]).bit_length() - 1
        op = Op(OpKind.MATK, tuple(n - 1 - b for b in target_bits), (), matrix=np.asarray(torch.as_tensor(matrix).numpy(), dtype=np.complex128))
        return oracle.run(Program(n, [op]), {}, psi.to(torch.complex128)).to(psi.dtype)

    monkeypatch.setattr(engine, "apply_dense", apply_dense)
    rng = np.random.default_rng(4)
    q, _ = np.linalg.qr(rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4)))
    noise = density.noise_superoperator((("AmplitudeDamping", (0.3,)), ("Depolarizing", (0.2,))))
    assert not np.allclose(noise @ noise.conj().T, np.eye(4))
    prog = Program(3, [Op(OpKind.MATK, (0, 1), (), matrix=q), Op(OpKind.MATK, (1, 2), (), matrix=noise), Op(OpKind.MATK, (2, 0), (), matrix=q.T.copy())])
    cp = engine.CompiledProgram(prog)
    psi0 = torch.tensor(rng.normal(size=(2, 8)) + 1j * rng.normal(size=(2, 8)))
    lam = torch.tensor(rng.normal(size=(2, 8)) + 1j * rng.normal(size=(2, 8)))
    psi = oracle.run(prog, {}, psi0)
    grads, psi_back, lam_back = engine._adjoint_walk(cp, {}, 2, torch.zeros(1, 2, dtype=torch.float64), psi.clone(), lam.clone())
    assert torch.allclose(psi_back, psi0, atol=1e-12)
    # <lam | M psi0> = <M† lam | psi0> for every psi0: compare through the inner product with a random probe
    probe = torch.tensor(rng.normal(size=(2, 8)) + 1j * rng.normal(size=(2, 8)))
    lhs = (lam.conj() * oracle.run(prog, {}, probe)).sum(-1)
    rhs = (lam_back.conj() * probe).sum(-1)
    assert torch.allclose(lhs, rhs, atol=1e-12)
    singular = Program(2, [Op(OpKind.MATK, (0, 1), (), matrix=np.diag([1.0, 0, 0, 1.0]).astype(np.complex128))])
    with pytest.raises(NotImplementedError, match="singular"):
        engine._adjoint_walk(engine.CompiledProgram(singular), {}, 1, torch.zeros(1, 1, dtype=torch.float64), psi0[:1, :4].clone(), lam[:1, :4].clone())
