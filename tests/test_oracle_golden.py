"""Pin the CPU oracle: every literal known-answer vector of the reference's own tests
(tests/golden/kat_reference.json cites file:line) plus internal consistency checks that need no
reference: adjoint gradients == torch autograd through the einsum forward, HamEvo == scipy expm,
sampler definition == a direct sequential inverse-CDF where no rounding ambiguity exists."""

from __future__ import annotations

import numpy as np
import pytest
import scipy.linalg
import torch

from oracle import statevec as oracle
from qadence_b200.program import PauliSum, PauliTerm, ProgramBuilder, hea_program, total_magnetization
from tests.kat import kat_program, kat_state, load_kats

KATS = load_kats()


@pytest.mark.parametrize("kat", KATS["run_zero_state"], ids=lambda k: k["ops"][0][0])
def test_kat_run_zero_state(kat: dict) -> None:
    wf = oracle.run(kat_program(kat["n"], kat["ops"]))
    assert torch.allclose(wf, kat_state(kat["state"]), atol=1e-14), kat["src"]


@pytest.mark.parametrize("kat", KATS["run_from_11"], ids=lambda k: k["ops"][0][0])
def test_kat_run_from_11(kat: dict) -> None:
    init = torch.tensor([[0, 0, 0, 1]], dtype=torch.complex128)
    wf = oracle.run(kat_program(kat["n"], kat["ops"]), state=init)
    assert torch.allclose(wf, kat_state(kat["state"]), atol=1e-14), kat["src"]


@pytest.mark.parametrize("kat", KATS["expect_Z0"], ids=lambda k: k["ops"][0][0])
def test_kat_expectation(kat: dict) -> None:
    z0 = PauliSum(1, [PauliTerm(1.0, (), ((0, "Z"),))])
    e = oracle.expectation(kat_program(kat["n"], kat["ops"]), [z0])
    assert e.shape == (1, 1) and abs(e.item() - kat["value"]) < 1e-14, kat["src"]


@pytest.mark.parametrize("kat", KATS["sample_counts_10_shots"], ids=lambda k: k["ops"][0][0])
def test_kat_sample(kat: dict) -> None:
    wf = oracle.run(kat_program(kat["n"], kat["ops"]))
    u = np.random.default_rng(0).random((1, 10))
    idx = oracle.sample_indices(oracle.probabilities(wf), u)
    assert oracle.counts_from_indices(idx, kat["n"]) == [kat["counts"]], kat["src"]


def test_kat_endianness() -> None:
    for case in KATS["endianness_X_on_3_qubits"]["cases"]:
        wf = oracle.run(ProgramBuilder(3).X(case["target"]).build())
        assert wf[0, case["index"]] == 1 and wf.abs().sum() == 1


def test_adjoint_equals_autograd() -> None:
    n, B = 4, 3
    b = ProgramBuilder(n)
    for q in range(n):
        b.RX(q, f"a{q}").RY(q, "shared").PHASE(q, f"p{q}").CRZ(q, (q + 1) % n, f"c{q}")
    b.scale("s")
    zz = PauliSum(n, [PauliTerm(0.7, (), ((0, "Z"), (1, "Z"))), PauliTerm(0.4, (), ((2, "X"),)), PauliTerm(-0.2, (), ((3, "Y"),))])
    b.hamevo(zz, "t")
    prog = b.build()
    obs = PauliSum(n, [PauliTerm(1.0, (), ((q, "Z"),)) for q in range(n)] + [PauliTerm(0.3, (), ((0, "X"), (2, "Y")))])
    g = torch.Generator().manual_seed(3)
    vals = {nm: torch.rand(B, generator=g, dtype=torch.float64).requires_grad_(True) for nm in prog.param_names}
    e = oracle.expectation(prog, [obs], vals)[:, 0]
    e.sum().backward()
    e2, grads = oracle.adjoint_expectation_and_grads(prog, obs, {k: v.detach() for k, v in vals.items()})
    assert torch.allclose(e.detach(), e2, atol=1e-13)
    for k, v in vals.items():
        assert torch.allclose(v.grad, grads[k], atol=1e-11), k


def test_hamevo_matches_scipy_expm() -> None:
    n = 3
    rng = np.random.default_rng(2)
    terms = [PauliTerm(0.9, (), ((0, "Z"), (2, "Z"))), PauliTerm(0.5, (), ((1, "X"),)), PauliTerm(-0.7, (), ((0, "Y"), (1, "Z")))]
    ps = PauliSum(n, terms)
    P = {"I": np.eye(2), "X": np.array([[0, 1], [1, 0]]), "Y": np.array([[0, -1j], [1j, 0]]), "Z": np.diag([1, -1])}
    H = np.zeros((8, 8), dtype=complex)
    for t in terms:
        mats = ["I"] * n
        for q, p in t.paulis:
            mats[q] = p
        full = np.array([[1.0]])
        for mname in mats:  # qubit 0 = MSB → leftmost kron factor
            full = np.kron(full, P[mname])
        H += t.const * full
    psi = rng.normal(size=(2, 8)) + 1j * rng.normal(size=(2, 8))
    tt = np.array([0.3, 1.7])
    prog = ProgramBuilder(n).hamevo(ps, "t").build()
    got = oracle.run(prog, {"t": torch.tensor(tt)}, torch.tensor(psi)).numpy()
    for b in range(2):
        want = scipy.linalg.expm(-1j * tt[b] * H) @ psi[b]
        assert np.allclose(got[b], want, atol=1e-12)


def test_sampler_definition_small() -> None:
    probs = np.array([[0.1, 0.2, 0.3, 0.4]])
    u = np.array([[0.05, 0.1, 0.29, 0.31, 0.59, 0.61, 0.999]])
    idx = oracle.sample_indices(probs, u)
    cdf = np.cumsum(probs[0])
    want = np.searchsorted(cdf, u[0] * cdf[-1], side="right")
    assert np.array_equal(idx[0], want)
    # multi-level tree (2^10 entries → 4 leaves of 256)
    rng = np.random.default_rng(0)
    p = rng.random((2, 1024))
    uu = rng.random((2, 64))
    idx = oracle.sample_indices(p, uu)
    for b in range(2):
        cdf = np.cumsum(p[b])
        ref = np.searchsorted(cdf, uu[b] * cdf[-1], side="right")
        assert np.abs(idx[b] - ref).max() <= 1  # identical up to rounding at cell boundaries


def test_hea_program_shape() -> None:
    prog = hea_program(20, 8)
    assert len(prog.ops) == 632 and len(prog.param_names) == 480  # 320 RX + 160 RY + 152 CNOT
    assert total_magnetization(20).is_diagonal


def test_oracle_exponential_is_exact_where_matrix_exp_is_not() -> None:
    """torch.linalg.matrix_exp (the reference's dense HamEvo) is off by ~2e-10 for small-norm complex128 matrices; the
    oracle's spectral exponential matches the closed form to rounding."""
    import torch

    from oracle import statevec as oracle

    x = torch.linspace(0.001, 0.3, 100, dtype=torch.float64)
    z = torch.tensor([[1, 0], [0, -1]], dtype=torch.complex128).expand(100, 2, 2)
    u = oracle.expm_i(z, x, -1j)
    assert (u[:, 0, 0] - torch.exp(-1j * x)).abs().max() < 1e-15 and (u[:, 1, 1] - torch.exp(1j * x)).abs().max() < 1e-15
    g = torch.tensor([[0.3, 0.2 - 0.1j], [0.2 + 0.1j, -0.7]], dtype=torch.complex128).expand(1, 2, 2)
    assert (oracle.expm_i(g, torch.tensor([1.7], dtype=torch.float64), -1j) - torch.linalg.matrix_exp(-1.7j * g)).abs().max() < 1e-9
