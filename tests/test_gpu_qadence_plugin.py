"""The REAL plugin on the GPU: qadence's own `QuantumModel` / `backend_factory` with `backend="b200sv"` driving the CUDA
kernels (no monkeypatching), mirroring the call sequence of `/root/reference/qadence/backends/pyqtorch/backend.py:255-281`
(`expectation` → `DifferentiableExpectation` → adjoint).  The qadence front-end comes from /root/reference in the CPU
container and from the copy under baseline/_ref on the GPU box (`__graft_entry__.install_reference_frontend`); the
test skips cleanly where neither exists.

The checker is the CPU oracle at USER level: the gate-parameter table is rebuilt from the model's variational
parameters and feature inputs with the torch restatement of the embedding (oracle/embedding.py), the converted program
runs through the einsum-per-gate oracle, and torch autograd gives the reference gradients w.r.t. the variational
parameters and the inputs."""

from __future__ import annotations

import os

import numpy as np
import pytest
import torch

from tests.conftest import enable_reference

pytestmark = [pytest.mark.gpu, pytest.mark.reference]
if not enable_reference():
    pytest.skip("qadence front-end not present", allow_module_level=True)

import qadence  # noqa: E402,F401
from qadence import QNN, QuantumCircuit, QuantumModel, chain, feature_map, hea, total_magnetization  # noqa: E402
from qadence.backends.api import backend_factory  # noqa: E402
from qadence.types import DiffMode  # noqa: E402

from oracle import embedding as oembed  # noqa: E402
from oracle import statevec as oracle  # noqa: E402


def _oracle_user_level(model, inputs: dict):
    """(expectation [B, n_obs], d/d vparams {name: grad}, d/d inputs {name: grad}) from the CPU oracle."""
    prog = model.embedding_fn.embed_program
    var = torch.stack([model._params[nm].detach().reshape(()).to(torch.float64) for nm in prog.var_names]).clone().requires_grad_(True) \
        if prog.var_names else torch.zeros(1, dtype=torch.float64, requires_grad=True)
    feats = [inputs[nm].detach().to(torch.float64).reshape(-1) for nm in prog.feat_names]
    batch = max([f.numel() for f in feats] + [1])
    feat = torch.stack([f.expand(batch) for f in feats]).clone().requires_grad_(True) if feats else torch.zeros(1, batch, dtype=torch.float64, requires_grad=True)
    table = oembed.evaluate(prog, var, feat, batch)
    values = {nm: table[i] for i, nm in enumerate(prog.names)}
    program = model._circuit.native.program
    obs = [o.native.obs for o in model._observable]
    e = oracle.expectation(program, obs, values)
    e.sum().backward()
    gv = {nm: var.grad[i] for i, nm in enumerate(prog.var_names)} if prog.var_names else {}
    gf = {nm: feat.grad[i] for i, nm in enumerate(prog.feat_names)}
    return e.detach(), gv, gf


def _check_model(model, inputs: dict, tol: float) -> None:
    leaf = {k: v.clone().requires_grad_(True) for k, v in inputs.items()}
    out = model.expectation(leaf)
    assert out.device.type == "cpu"  # results come back on the device of the inputs (drop-in for a CPU QuantumModel)
    out.sum().backward()
    want_e, want_gv, want_gf = _oracle_user_level(model, inputs)
    assert out.shape == want_e.shape
    assert (out.detach().double() - want_e).abs().max().item() <= tol
    for nm, g in want_gv.items():
        got = model._params[nm].grad
        assert got is not None, nm
        assert abs(float(got.reshape(())) - float(g)) <= 50 * tol, nm
    for nm, g in want_gf.items():
        assert (leaf[nm].grad.double() - g).abs().max().item() <= 50 * tol, nm


@pytest.mark.parametrize("diff_mode", [DiffMode.ADJOINT, DiffMode.AD, DiffMode.GPSR])
def test_cfg1_qnn_through_the_plugin(diff_mode) -> None:
    """BASELINE.json configs[0]: QNN hea(4, 2) expectation of total_magnetization, batch 16."""
    n = 4
    torch.manual_seed(0)
    circ = QuantumCircuit(n, chain(feature_map(n, param="x"), hea(n, 2)))
    model = QuantumModel(circ, total_magnetization(n), backend="b200sv", diff_mode=diff_mode)
    assert model.backend.backend.name == "b200sv"
    x = torch.rand(16, dtype=torch.float64, generator=torch.Generator().manual_seed(1))
    _check_model(model, {"x": x}, 1e-10)


def test_cfg4_shaped_qnn_adjoint_through_the_plugin() -> None:
    """configs[3] shape at full width: feature map + hea(16, 12) (772 gates), 2 samples, adjoint through qadence's own
    `DifferentiableBackend` — the lean TMA sweeps are what executes (n = 16 > tile)."""
    n = 16
    torch.manual_seed(3)
    circ = QuantumCircuit(n, chain(feature_map(n, param="x"), hea(n, 12)))
    model = QNN(circ, total_magnetization(n), inputs=["x"], backend="b200sv", diff_mode=DiffMode.ADJOINT)
    x = torch.tensor([0.31, -0.77], dtype=torch.float64)
    _check_model(model, {"x": x}, 1e-10)


def test_complex64_configuration_and_run_sample() -> None:
    n = 6
    torch.manual_seed(5)
    circ = QuantumCircuit(n, chain(feature_map(n, param="x"), hea(n, 2)))
    model = QuantumModel(circ, total_magnetization(n), backend="b200sv", diff_mode=DiffMode.ADJOINT, configuration={"dtype": "complex64"})
    x = torch.rand(5, dtype=torch.float64, generator=torch.Generator().manual_seed(2))
    _check_model(model, {"x": x}, 1e-5)
    model128 = QuantumModel(circ, total_magnetization(n), backend="b200sv", diff_mode=DiffMode.ADJOINT)
    psi = model128.run({"x": x})
    prog = model128.embedding_fn.embed_program
    var = torch.stack([model128._params[nm].detach().reshape(()) for nm in prog.var_names])
    table = oembed.evaluate(prog, var, x.reshape(1, -1), 5)
    want = oracle.run(model128._circuit.native.program, {nm: table[i] for i, nm in enumerate(prog.names)})
    assert (psi.cpu() - want).abs().max().item() <= 1e-10
    counts = model128.sample({"x": x}, n_shots=200)
    assert len(counts) == 5 and all(sum(c.values()) == 200 and all(len(k) == n for k in c) for c in counts)


def test_backend_factory_call_sequence() -> None:
    """The sequence `QuantumModel` performs, spelled out (backend_factory → convert → embedding_fn → expectation)."""
    n = 5
    circ = QuantumCircuit(n, chain(feature_map(n, param="x"), hea(n, 1)))
    bk = backend_factory("b200sv", diff_mode="adjoint")
    conv = bk.convert(circ, total_magnetization(n))
    x = torch.linspace(-0.5, 0.5, 3, dtype=torch.float64)
    params = {k: (v.detach().clone().requires_grad_(True) if v.requires_grad else v) for k, v in conv.params.items()}
    e = bk.expectation(conv.circuit, conv.observable, conv.embedding_fn(params, {"x": x}))
    assert e.shape == (3, 1)
    e.sum().backward()
    assert any(p.grad is not None and p.grad.abs().sum() > 0 for p in params.values() if p.requires_grad)


def test_lindblad_noise_operators_and_dropout_through_the_plugin() -> None:
    """`HamEvo(..., noise_operators=[...])` (tests/backends/pyq/test_time_dependent_generator.py:68-126) returns the density
    matrix of the integrated master equation; expectation = Tr(Oρ).  Quantum dropout: the training-mode circuit equals
    the oracle on the SAME dropped sub-program, evaluation mode equals the full circuit."""
    from tests.test_lindblad import integrate_master_equation

    X2, Z2, I2 = np.array([[0, 1], [1, 0]], dtype=complex), np.diag([1.0 + 0j, -1]), np.eye(2)
    gen = 0.9 * qadence.X(0) + 0.6 * qadence.Z(0) * qadence.Z(1) + 0.3 * qadence.Y(1)
    block = chain(qadence.RX(0, 0.4), qadence.HamEvo(gen, 1.2, noise_operators=[0.7 * qadence.X(1), 0.4 * qadence.Z(0)]))
    model = QuantumModel(QuantumCircuit(2, block), total_magnetization(2), backend="b200sv", diff_mode=DiffMode.GPSR)
    rho = model.run({})
    assert type(rho).__name__ == "DensityMatrix"
    H = 0.9 * np.kron(X2, I2) + 0.6 * np.kron(Z2, Z2) + 0.3 * np.kron(I2, np.array([[0, -1j], [1j, 0]]))
    psi = oracle.run(model.backend.backend.circuit(QuantumCircuit(2, qadence.RX(0, 0.4))).native.program, {})[0].numpy()
    want = integrate_master_equation(lambda t: H, [0.7 * np.kron(I2, X2), 0.4 * np.kron(Z2, I2)], np.outer(psi, psi.conj()), 1.2)
    assert np.abs(rho[0].cpu().numpy() - want).max() < 1e-9
    e = model.expectation({})
    assert abs(float(e) - np.trace((np.kron(Z2, I2) + np.kron(I2, Z2)) @ want).real) < 1e-9

    n = 4
    circ = QuantumCircuit(n, chain(feature_map(n, param="x"), hea(n, 2)))
    model = QuantumModel(circ, total_magnetization(n), backend="b200sv", diff_mode=DiffMode.ADJOINT, configuration={"dropout_probability": 0.5})
    x = {"x": torch.linspace(-0.5, 0.5, 3, dtype=torch.float64)}
    nat = model._circuit.native
    vals = model.embedding_fn(model._params, x)
    host = {k: torch.as_tensor(v).detach().cpu() for k, v in vals.items()}
    model.train()
    torch.manual_seed(3)
    sub = nat.compiled_for_call(True, vals)
    torch.manual_seed(3)
    e_train = model.expectation(x)
    assert (e_train.detach().cpu() - oracle.expectation(sub.program, [total_magnetization_sum(n)], host)).abs().max() < 1e-10
    assert len(sub.program.ops) < len(nat.program.ops)
    e_train.sum().backward()
    assert any(p.grad is not None for p in model.parameters())
    model.eval()
    e_eval = model.expectation(x)
    assert (e_eval.detach().cpu() - oracle.expectation(nat.program, [total_magnetization_sum(n)], host)).abs().max() < 1e-10


def total_magnetization_sum(n: int):
    from qadence_b200.program import total_magnetization as tm

    return tm(n)


def test_replace_mode_in_a_subprocess() -> None:
    """`B200SV_REPLACE_PYQTORCH=1`: `backend="pyqtorch"` (qadence's default) builds the b200sv backend and runs on the GPU."""
    import subprocess
    import sys
    from pathlib import Path

    from tests.conftest import REFERENCE, ROOT

    code = f"""
import sys
sys.path[:0] = [{str(ROOT)!r}, {str(ROOT / 'tests' / '_shims')!r}, {str(REFERENCE)!r}]
import torch
from qadence import QuantumCircuit, QuantumModel, chain, feature_map, hea, total_magnetization
n = 4
torch.manual_seed(0)
m = QuantumModel(QuantumCircuit(n, chain(feature_map(n, param="x"), hea(n, 2))), total_magnetization(n), diff_mode="adjoint")
assert m.backend.backend.name == "pyqtorch" and type(m.backend.backend).__module__.startswith("qadence_b200"), type(m.backend.backend).__module__
x = torch.tensor([0.2, 0.5], requires_grad=True)
e = m.expectation({{"x": x}})
e.sum().backward()
print("REPLACE_OK", e.flatten().tolist(), x.grad.tolist())
"""
    env = {**os.environ, "B200SV_REPLACE_PYQTORCH": "1", "QADENCE_LOG_LEVEL": "ERROR"}
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0 and "REPLACE_OK" in out.stdout, out.stderr[-2000:]
