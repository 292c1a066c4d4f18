"""Density-matrix / noise path on the GPU: the vectorised programs of qadence_b200/density.py through the C-ABI,
against the CPU oracle running the same transformation and against the independent dense Kraus evolution."""

from __future__ import annotations

import numpy as np
import pytest
import torch

from oracle import statevec as oracle
from qadence_b200 import density, engine
from qadence_b200.program import PauliSum, PauliTerm, ProgramBuilder, hea_program, total_magnetization
from tests.test_density import PROTOCOLS, dense_density, random_noisy_program, run_vectorised

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("seed", range(8))
def test_random_noisy_programs(seed: int) -> None:
    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(1, 5))
    prog, values = random_noisy_program(rng, n, int(rng.integers(4, 12)))
    cd = density.CompiledDensityProgram(prog)
    rho = density.run(cd, values, device=DEV)
    assert type(rho).__name__ == "DensityMatrix" and rho.shape[1:] == (2**n, 2**n)
    want = run_vectorised(prog, values)
    assert (rho.as_subclass(torch.Tensor).cpu() - want).abs().max() < 1e-10
    if n <= 3:
        b = rho.shape[0] - 1
        rho0 = np.zeros((2**n, 2**n), dtype=np.complex128)
        rho0[0, 0] = 1
        assert np.abs(rho[b].as_subclass(torch.Tensor).cpu().numpy() - dense_density(prog, values, rho0, b)).max() < 1e-10
    obs = PauliSum(n, [PauliTerm(0.7 + 0j, (), ((0, "Z"),)), PauliTerm(-0.4 + 0j, (), ((n - 1, "X"),)), PauliTerm(0.3 + 0j, (), ((0, "Y"),))])
    ex = density.expectation(cd, [engine.CompiledObservable(density.lift_observable(obs, n))], values, device=DEV).cpu()
    o = np.zeros((2**n, 2**n), dtype=np.complex128)
    eye = torch.eye(2**n, dtype=torch.complex128)
    o = oracle.from_pyq_shape(oracle.paulisum_apply(obs, oracle.to_pyq_shape(eye, n), {})).numpy().T
    want_e = np.einsum("ij,bji->b", o, want.numpy()).real
    assert np.abs(ex[:, 0].numpy() - want_e).max() < 1e-10


def test_noisy_hea_trace_purity_and_complex64() -> None:
    n = 6
    prog = hea_program(n, 2)
    for k, op in enumerate(prog.ops):
        name, probs = PROTOCOLS[k % len(PROTOCOLS)]
        op.noise = ((name, probs),)
    rng = np.random.default_rng(1)
    values = {nm: torch.tensor(rng.uniform(0, 6, size=3)) for nm in prog.param_names}
    cd = density.CompiledDensityProgram(prog)
    rho = density.run(cd, values, device=DEV).as_subclass(torch.Tensor)
    tr = rho.diagonal(dim1=1, dim2=2).sum(-1)
    assert (tr - 1).abs().max() < 1e-12
    assert (rho - rho.conj().transpose(1, 2)).abs().max() < 1e-12
    purity = torch.einsum("bij,bji->b", rho, rho).real
    assert (purity < 0.9).all() and (purity > 1 / 2**n).all()
    want = run_vectorised(prog, values)
    assert (rho.cpu() - want).abs().max() < 1e-10
    rho32 = density.run(cd, values, dtype=torch.complex64, device=DEV).as_subclass(torch.Tensor)
    assert rho32.dtype == torch.complex64 and (rho32.cpu().to(torch.complex128) - want).abs().max() < 1e-5
    ex = density.expectation(cd, [engine.CompiledObservable(density.lift_observable(total_magnetization(n), n))], values, device=DEV)
    z = torch.tensor([[n - 2 * bin(i).count("1") for i in range(2**n)]], dtype=torch.float64)
    assert (ex[:, 0].cpu() - (want.diagonal(dim1=1, dim2=2).real * z).sum(-1)).abs().max() < 1e-10


def test_readout_confusion_and_sampling() -> None:
    n = 4
    prog = ProgramBuilder(n).H(0).CNOT(0, 1).X(3).noise("BitFlip", 0.2).build()
    cd = density.CompiledDensityProgram(prog)
    probs = density.probabilities(density.run(cd, {}, device=DEV))
    assert abs(float(probs.sum()) - 1) < 1e-12
    ro = density.ReadoutNoise(n, error_probability=[0.1, 0.0, 0.05, 0.2])
    noisy = ro.apply(probs, device=DEV)
    want = oracle.run(ro.program(), {}, probs.cpu().to(torch.complex128)).real
    assert (noisy.cpu() - want).abs().max() < 1e-14 and abs(float(noisy.sum()) - 1) < 1e-12
    u = torch.rand(1, 20000, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    counts = density.sample_probabilities(noisy, 20000, uniforms=u)[0]
    assert sum(counts.values()) == 20000
    for idx in range(2**n):
        assert abs(counts.get(format(idx, f"0{n}b"), 0) / 20000 - float(want[0, idx])) < 0.015
