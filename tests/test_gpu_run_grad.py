"""Differentiable `engine.run` on the GPU (`_RunState`: torch's cotangent of the output state fed to the adjoint sweeps)
against torch autograd through the CPU oracle — what pyqtorch's autograd-differentiable `run` gives the reference
(/root/reference/tests/qadence/test_overlap.py:304-337 trains through it)."""

from __future__ import annotations

import numpy as np
import pytest
import torch

from oracle import statevec as oracle
from qadence_b200 import engine
from qadence_b200.program import PauliSum, PauliTerm, ProgramBuilder

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("dtype,tol", [(torch.complex128, 1e-10), (torch.complex64, 2e-5)])
def test_run_gradients_match_autograd_through_the_oracle(dtype: torch.dtype, tol: float) -> None:
    rng = np.random.default_rng(3)
    n, B = 7, 3
    b = ProgramBuilder(n)
    for q in range(n):
        b.RX(q, f"a{q}").RY(q, "shared").PHASE(q, f"p{q}").H(q)
    for q in range(n - 1):
        b.CNOT(q, q + 1).CRZ(q, q + 1, f"c{q}").CPHASE(q + 1, q, "shared")
    b.scale("s")
    b.hamevo(PauliSum(n, [PauliTerm(0.7 + 0j, (), ((0, "X"), (1, "X"))), PauliTerm(-0.4 + 0j, (), ((2, "Z"),)), PauliTerm(0.3 + 0j, (), ((3, "Y"), (5, "Z")))]), "t")
    for q in range(n):
        b.RZ(q, f"z{q}").RX(q, 0.3)
    prog = b.build()
    cp = engine.CompiledProgram(prog)
    base = {nm: rng.uniform(0.2, 1.5, size=B if k % 3 else 1) for k, nm in enumerate(prog.param_names)}
    psi0 = rng.normal(size=(1, 2**n)) + 1j * rng.normal(size=(1, 2**n))
    psi0 /= np.linalg.norm(psi0)
    target = torch.tensor(rng.normal(size=(B, 2**n)) + 1j * rng.normal(size=(B, 2**n)))

    def loss_of(psi: torch.Tensor) -> torch.Tensor:
        psi = psi.to(torch.complex128).cpu()
        return ((psi - target).abs() ** 2).sum() + (psi.imag * target.real).sum()

    vals = {k: torch.tensor(v, requires_grad=True) for k, v in base.items()}
    st = torch.tensor(psi0, requires_grad=True)
    out = engine.run(cp, vals, st, dtype=dtype, device=DEV)
    assert out.requires_grad and out.dtype == dtype and out.shape == (B, 2**n)
    loss_of(out).backward()
    ref_vals = {k: torch.tensor(v, requires_grad=True) for k, v in base.items()}
    ref_st = torch.tensor(psi0, requires_grad=True)
    ref_out = oracle.run(prog, ref_vals, ref_st.expand(B, -1))
    assert (out.detach().cpu().to(torch.complex128) - ref_out.detach()).abs().max() < tol
    loss_of(ref_out).backward()
    for k in vals:
        assert vals[k].grad is not None and vals[k].grad.shape == ref_vals[k].grad.shape, k
        assert (vals[k].grad - ref_vals[k].grad).abs().max() < tol * 50, (k, vals[k].grad, ref_vals[k].grad)
    assert st.grad.shape == (1, 2**n) and (st.grad - ref_st.grad).abs().max() < tol * 50
    with torch.no_grad():
        plain = engine.run(cp, vals, st, dtype=dtype, device=DEV)
    assert not plain.requires_grad and torch.equal(plain, out.detach())
