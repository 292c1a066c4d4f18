"""Gradients of noisy expectations on the GPU: `density.expectation` in grad mode = a linear functional of the differentiable
`engine.run` of the vectorised program (adjoint walk with λ = vec(O), noise superoperators un-applied with their inverse),
against torch autograd through the CPU oracle running the same vectorised program.  (Sorted last on purpose: the composition
was validated piecewise — `_RunState` and `apply_dense` on the GPU, the algebra on CPU — when the round's GPU budget ran out.)"""

from __future__ import annotations

import numpy as np
import pytest
import torch

from oracle import statevec as oracle
from qadence_b200 import density, engine
from qadence_b200.program import PauliSum, PauliTerm, ProgramBuilder

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_noisy_expectation_gradients() -> None:
    rng = np.random.default_rng(9)
    n, B = 3, 2
    b = ProgramBuilder(n)
    b.RX(0, "a").noise("Depolarizing", 0.1).RY(1, "b").noise("AmplitudeDamping", 0.2).CNOT(0, 1).noise("BitFlip", 0.05)
    b.RZ(2, "a").CRX(1, 2, "c").noise("PhaseDamping", 0.15).H(0).PHASE(2, "b").noise("GeneralizedAmplitudeDamping", 0.6, 0.3)
    prog = b.build()
    obs = [PauliSum(n, [PauliTerm(1.0 + 0j, (), ((q, "Z"),)) for q in range(n)]), PauliSum(n, [PauliTerm(0.5 + 0j, (), ((0, "X"), (2, "Y")))])]
    cd = density.CompiledDensityProgram(prog)
    cobs = [engine.CompiledObservable(density.lift_observable(o, n)) for o in obs]
    base = {"a": rng.uniform(0.2, 1.5, size=B), "b": rng.uniform(0.2, 1.5, size=1), "c": rng.uniform(0.2, 1.5, size=B)}
    w = torch.tensor([[0.7, -1.3]])

    vals = {k: torch.tensor(v, requires_grad=True) for k, v in base.items()}
    e = density.expectation(cd, cobs, vals, device=DEV)
    assert e.requires_grad and e.shape == (B, 2)
    (e.cpu() * w).sum().backward()

    ref = {k: torch.tensor(v, requires_grad=True) for k, v in base.items()}
    vprog, needs = density.vectorise(prog)
    v = oracle.run(vprog, density.extend_values(ref, needs), None)
    N = 2**n
    cols = []
    for o in obs:
        eye = torch.eye(N, dtype=torch.complex128)
        om = oracle.from_pyq_shape(oracle.paulisum_apply(o, oracle.to_pyq_shape(eye, n), {})).T  # dense O
        cols.append(torch.einsum("ij,bji->b", om, v.reshape(-1, N, N)).real)
    e_ref = torch.stack(cols, dim=1)
    assert (e.detach().cpu() - e_ref.detach()).abs().max() < 1e-10
    (e_ref * w).sum().backward()
    for k in vals:
        assert vals[k].grad is not None and (vals[k].grad - ref[k].grad).abs().max() < 1e-9, (k, vals[k].grad, ref[k].grad)
    with torch.no_grad():
        plain = density.expectation(cd, cobs, vals, device=DEV)
    assert not plain.requires_grad and (plain.cpu() - e.detach().cpu()).abs().max() < 1e-12
