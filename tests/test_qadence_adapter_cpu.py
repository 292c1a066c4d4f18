"""The qadence-facing adapter (registration, conversion, Backend API semantics) on CPU.

Needs the reference front-end (`/root/reference`, through tests/_shims) → marker `reference`.
There is no GPU here and the product has no CPU fallback, so where a test must push numbers
through `Backend.run/expectation/sample` the C-ABI-backed engine functions are monkeypatched with
the CPU oracle (a test double living entirely in tests/); conversion itself is checked against the
reference's own dense oracle `block_to_tensor`, like /root/reference/tests/qadence/test_matrices.py.
"""

from __future__ import annotations

from collections import Counter

import numpy as np
import pytest
import torch

from tests.conftest import enable_reference

pytestmark = pytest.mark.reference
if not enable_reference():
    pytest.skip("/root/reference not present", allow_module_level=True)

import qadence  # noqa: E402
import qadence.states as qstates  # noqa: E402
from qadence import (  # noqa: E402
    CNOT, CPHASE, CRX, CRZ, RX, RY, RZ, H, HamEvo, I, N, QuantumCircuit, QuantumModel, X, Y, Z, add, chain, feature_map, hea, kron,
    qft, total_magnetization,
)
from qadence.backends.api import backend_factory  # noqa: E402
from qadence.blocks.block_to_tensor import block_to_tensor  # noqa: E402
from qadence.types import BackendName, DiffMode, Endianness, Engine  # noqa: E402

from oracle import statevec as oracle  # noqa: E402
from qadence_b200 import cabi, engine  # noqa: E402


def _product_state(bitstring: str, batch_size: int = 1, **_: object) -> torch.Tensor:
    s = torch.zeros(batch_size, 2 ** len(bitstring), dtype=torch.cdouble)
    s[:, int(bitstring, 2)] = 1.0
    return s


qstates.product_state = _product_state


@pytest.fixture
def oracle_device(monkeypatch: pytest.MonkeyPatch) -> None:
    """Replace the GPU entry points of `qadence_b200.engine` by the CPU oracle (test double)."""

    def run(cp, values=None, state=None, dtype=torch.complex128, device=None):
        # differentiable like the product's `engine.run` (tests/test_gpu_run_grad.py): autograd through the oracle's torch ops
        vals = {k: (torch.as_tensor(v) if torch.is_grad_enabled() else torch.as_tensor(v).detach()) for k, v in (values or {}).items()}
        return oracle.run(cp.program, vals, state).to(dtype)

    def expectation(cp, observables, values=None, state=None, dtype=torch.complex128, device=None):
        out = oracle.expectation(cp.program, [o.obs for o in observables], values or {}, state)
        return out.to(torch.float64 if dtype == torch.complex128 else torch.float32)

    def sample(cp, values=None, n_shots=1, state=None, dtype=torch.complex128, device=None, uniforms=None, little_endian=False):
        psi = oracle.run(cp.program, values or {}, state)
        u = np.random.default_rng(0).random((psi.shape[0], n_shots)) if uniforms is None else uniforms.numpy()
        idx = oracle.sample_indices(oracle.probabilities(psi), u)
        return [Counter({(k[::-1] if little_endian else k): v for k, v in d.items()}) for d in oracle.counts_from_indices(idx, cp.n_qubits)]

    def apply_observable(cob, psi, values):
        if cob.pauli is not None:
            return oracle.from_pyq_shape(oracle.paulisum_apply(cob.obs, oracle.to_pyq_shape(psi, cob.n_qubits), values))
        return oracle.run(cob.obs, values, psi)

    def sample_indices(psi, uniforms):
        return torch.as_tensor(oracle.sample_indices(oracle.probabilities(psi), uniforms.numpy()))

    monkeypatch.setattr(engine, "run", run)
    monkeypatch.setattr(engine, "expectation", expectation)
    monkeypatch.setattr(engine, "sample", sample)
    monkeypatch.setattr(engine, "apply_observable", apply_observable)
    monkeypatch.setattr(engine, "sample_indices", sample_indices)
    monkeypatch.setattr(engine, "bit_reverse", lambda psi: oracle.invert_endianness_state(psi, int(psi.shape[1]).bit_length() - 1))


# ------------------------------------------------------------------------------------------ registration
def test_backend_is_registered_without_touching_qadence() -> None:
    assert "b200sv" in BackendName.list() and BackendName("b200sv") == "b200sv"
    assert "b200sv" in Engine.list()
    for mode in (DiffMode.AD, DiffMode.ADJOINT, DiffMode.GPSR):
        bk = backend_factory("b200sv", diff_mode=mode)
        assert type(bk).__name__ == "DifferentiableBackend" and bk.backend.name == "b200sv"
        assert bk.backend.config._use_gate_params is True
    plain = backend_factory("b200sv", diff_mode=None)
    assert type(plain).__name__ == "Backend"
    # dict configuration resolves through qadence.backends.b200sv.config (import_config)
    bk = backend_factory("b200sv", diff_mode="adjoint", configuration={"dtype": "complex64"})
    assert bk.backend.config.dtype == "complex64"
    from qadence.extensions import supported_gates

    names = {g.__name__ for g in supported_gates("b200sv")}
    assert {"RX", "CNOT", "HamEvo", "Toffoli", "MCPHASE", "Projector"} <= names and "TDagger" not in names
    from dataclasses import asdict

    assert "dtype" in asdict(bk.backend.config)  # QuantumModel._to_dict (model.py:487)


def test_no_cpu_fallback() -> None:
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    bk = backend_factory("b200sv", diff_mode=None)
    conv = bk.convert(QuantumCircuit(2, X(0)))
    with pytest.raises(cabi.B200svError):
        bk.run(conv.circuit, {})
    conv = bk.convert(QuantumCircuit(2, RX(0, "x")))  # the device-side parameter embedding has no host fallback either
    with pytest.raises(cabi.B200svError):
        conv.embedding_fn(conv.params, {"x": torch.tensor([0.1])})


# ------------------------------------------------------------------------------------------ conversion
BLOCKS = [
    chain(X(0), Y(1), Z(2), H(0)),
    chain(RX(0, 0.3), RY(1, "y"), RZ(2, "y"), CNOT(0, 2), CRX(1, 0, "y"), CPHASE(2, 1, 0.4)),
    hea(3, 2),
    chain(feature_map(3, param="y"), hea(3, 1)),
    qft(3),
    chain(HamEvo(add(0.7 * kron(Z(0), Z(1)), 0.3 * X(2), 1.1 * N(1)), 0.9), H(1)),
    chain(kron(H(0), H(1), H(2)), add(kron(X(0), I(1)), 0.5 * Z(2))),
    chain(2.0 * X(0), "y" * CRZ(0, 1, 0.2)),
]


@pytest.mark.parametrize("block", BLOCKS, ids=lambda b: type(b).__name__ + str(len(str(b)) % 97))
def test_converted_program_equals_dense_oracle(block) -> None:
    """Mirror of tests/qadence/test_matrices.py:68-72,582-771 for the b200sv converter."""
    n = 3
    bk = backend_factory("b200sv", diff_mode=None)
    conv = bk.convert(QuantumCircuit(n, block))
    inputs = {"y": torch.tensor([0.41])}
    vals = conv.embedding_fn(conv.params, inputs)
    rng = np.random.default_rng(1)
    psi = torch.tensor(rng.normal(size=(1, 8)) + 1j * rng.normal(size=(1, 8)))
    user = {**{k: v for k, v in conv.params.items() if not k.startswith("fix_")}, **inputs}
    dense = block_to_tensor(block, values=user, qubit_support=tuple(range(n)))
    want = torch.einsum("bij,bj->bi", dense, psi)
    got = oracle.run(conv.circuit.native.program, {k: torch.as_tensor(v).detach() for k, v in vals.items()}, psi)
    assert (got - want.detach()).abs().max().item() < 1e-12


def test_observable_becomes_pauli_sum() -> None:
    from qadence_b200.program import PauliSum

    bk = backend_factory("b200sv", diff_mode=None)
    obs = add(0.5 * X(0), kron(Y(1), Z(2)), 1.2 * N(2), "w" * Z(0))
    conv = bk.convert(QuantumCircuit(3, hea(3, 1)), obs)
    ps = conv.observable[0].native.obs
    assert isinstance(ps, PauliSum) and len(ps.param_names) == 1
    zz = conv = bk.convert(QuantumCircuit(3, hea(3, 1)), total_magnetization(3)).observable[0].native.obs
    assert zz.is_diagonal and len(zz.terms) == 3


def test_unlowered_analog_block_error() -> None:
    from qadence import AnalogRX
    from qadence_b200.qadence_backend.convert_ops import convert_block

    with pytest.raises(NotImplementedError, match="device_specs"):
        convert_block(AnalogRX(0.3), 2, backend_factory("b200sv", diff_mode=None).config)


def test_cfg2_program_matches_standalone_builder() -> None:
    """bench.py's standalone hea_program(20, 8) is the circuit qadence's hea(20, 8) converts to."""
    from qadence_b200.program import hea_program

    bk = backend_factory("b200sv", diff_mode="adjoint")
    conv = bk.convert(QuantumCircuit(20, hea(20, 8)))
    ours = hea_program(20, 8)
    theirs = conv.circuit.native.program
    assert len(theirs.ops) == len(ours.ops) == 632
    uuid_to_name = {}
    vals = conv.embedding_fn(conv.params, {})
    assert len(vals) == 480
    for a, b in zip(theirs.ops, ours.ops):
        assert (a.kind, a.targets, a.controls) == (b.kind, b.targets, b.controls)
        if isinstance(a.param, str):
            uuid_to_name[a.param] = b.param
    # gate k of the reference uses theta_k of the standalone builder
    for u, nm in uuid_to_name.items():
        assert torch.allclose(vals[u], conv.params[nm])


# ------------------------------------------------------------------------------------------ Backend API semantics
def test_state_validation_errors(oracle_device) -> None:
    bk = backend_factory("b200sv", diff_mode=None)
    conv = bk.convert(QuantumCircuit(2, kron(X(0), X(1))))
    with pytest.raises(ValueError):  # test_quantum_pyq.py:164-169
        bk.run(conv.circuit, state=torch.tensor([1.0, 0.0], dtype=torch.complex128))
    with pytest.raises(TypeError):
        bk.run(conv.circuit, state=torch.zeros(1, 4, dtype=torch.float64))
    with pytest.raises(NotImplementedError):
        bk.assign_parameters(conv.circuit, {})
    # NB the reference's `is_pyq_shape` compares a torch.Size with a list (backends/utils.py:150-151) and is
    # therefore always False: pyqtorch-shaped states are rejected by `validate_state` there too.
    pyq_shaped = torch.zeros(2, 2, 1, dtype=torch.complex128)
    with pytest.raises(ValueError):
        bk.run(conv.circuit, state=pyq_shaped)
    init = torch.tensor([[0, 0, 0, 1]], dtype=torch.complex128)
    out = bk.run(conv.circuit, state=init)
    assert out.shape == (1, 4) and out[0, 0] == 1


def test_run_endianness_and_sample(oracle_device) -> None:
    bk = backend_factory("b200sv", diff_mode=None)
    conv = bk.convert(QuantumCircuit(3, X(0)))  # tests/backends/test_endianness.py:113-138
    assert bk.run(conv.circuit)[0, 4] == 1
    assert bk.run(conv.circuit, endianness=Endianness.LITTLE)[0, 1] == 1
    assert bk.sample(conv.circuit, n_shots=100) == [Counter({"100": 100})]
    assert bk.sample(conv.circuit, n_shots=100, endianness=Endianness.LITTLE) == [Counter({"001": 100})]


def test_quantum_model_expectation_and_gradients(oracle_device) -> None:
    """QuantumModel / QNN plumbing: values → embedding_fn → DifferentiableBackend → [B, n_obs], and
    the autograd edge back to the nn.Parameters (configs[0] shape)."""
    torch.manual_seed(0)
    circ = QuantumCircuit(4, feature_map(4), hea(4, 2))
    obs = [total_magnetization(4), kron(X(0), Y(1))]
    model = QuantumModel(circ, obs, backend="b200sv", diff_mode=DiffMode.ADJOINT)
    x = torch.linspace(-0.9, 0.9, 16)
    out = model.expectation({"phi": x})
    assert out.shape == (16, 2)
    out[:, 0].sum().backward()
    grads = {k: p.grad.clone() for k, p in model._params.items() if p.requires_grad}
    assert len(grads) == 24 and all(g is not None for g in grads.values())
    user = {k: p.detach().clone().requires_grad_(p.requires_grad) for k, p in model._params.items()}
    dense = block_to_tensor(circ.block, values={**user, "phi": x}, qubit_support=(0, 1, 2, 3))
    psi = dense[:, :, 0]
    om = block_to_tensor(obs[0], qubit_support=(0, 1, 2, 3))
    e = torch.einsum("bi,ij,bj->b", psi.conj(), om[0], psi).real
    assert torch.allclose(out[:, 0].detach(), e.detach(), atol=1e-12)
    e.sum().backward()
    for k, g in grads.items():
        assert torch.allclose(g, user[k].grad, atol=1e-10), k


def test_gpsr_equals_adjoint(oracle_device) -> None:
    """tests/engines/test_torch.py:111-134 — adjoint == gpsr on the same circuit, through the plugin."""
    torch.manual_seed(1)
    circ = QuantumCircuit(2, chain(RX(0, "x"), RY(1, "t0"), CNOT(0, 1), RX(1, "t1")))
    obs = total_magnetization(2)
    vals = {"x": torch.tensor([0.3, 0.8])}
    grads = {}
    for mode in (DiffMode.ADJOINT, DiffMode.GPSR):
        torch.manual_seed(1)
        model = QuantumModel(circ, obs, backend="b200sv", diff_mode=mode)
        for p in model._params.values():
            if p.requires_grad:
                p.data.fill_(0.7)
        out = model.expectation(vals)
        out.sum().backward()
        grads[mode] = torch.stack([p.grad.reshape(()) for k, p in sorted(model._params.items()) if p.requires_grad and p.grad is not None])
        assert grads[mode].numel() >= 2
    assert torch.allclose(grads[DiffMode.ADJOINT], grads[DiffMode.GPSR], atol=1e-6)


def test_parametric_observable_only_in_ad_mode() -> None:
    bk = backend_factory("b200sv", diff_mode=DiffMode.ADJOINT)
    with pytest.raises(ValueError):
        bk.observable("w" * Z(0), 1)


def test_qnn_training_step_through_the_plugin(oracle_device) -> None:
    """configs[0] end to end at the ML layer: `QNN(hea(4, 2) + feature map, total_magnetization)` built by qadence's own
    ml_tools on the b200sv backend, one Adam step on an MSE loss (ml_tools/models.py:107-475, optimize_step.py:44-54)."""
    from qadence.ml_tools.models import QNN

    torch.manual_seed(0)
    circ = QuantumCircuit(4, feature_map(4), hea(4, 2))
    model = QNN(circ, total_magnetization(4), backend="b200sv", diff_mode=DiffMode.ADJOINT)
    x = torch.linspace(-0.9, 0.9, 16).reshape(-1, 1)
    y = x**2
    opt = torch.optim.Adam(model.parameters(), lr=0.05)
    losses = []
    for _ in range(3):
        opt.zero_grad()
        out = model(x)
        assert out.shape == (16, 1)
        loss = torch.nn.functional.mse_loss(out, y.to(out.dtype))
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < losses[0]
    assert all(p.grad is not None for p in model.parameters() if p.requires_grad)


def test_time_dependent_hamevo_conversion() -> None:
    """The reference's own time-dependent generator (tests/conftest.py:261-266) becomes a symbolic-coefficient Pauli sum with
    `duration` as the evolution parameter and `n_steps_hevo` steps (backends/pyqtorch/convert_ops.py:227-231,258-268)."""
    import sympy
    from qadence.parameters import Parameter, TimeParameter

    from tests.test_time_dependent import FY, OMEGA, X as Xm, Y as Ym, I2, ode_solution

    t, x = TimeParameter("t"), Parameter("x", trainable=False)
    gen = OMEGA * (sympy.sin(FY * t) * qadence.X(0) + x * (t**2) * qadence.Y(1))
    bk = backend_factory("b200sv", diff_mode=None, configuration={"n_steps_hevo": 500})
    conv = bk.convert(QuantumCircuit(2, qadence.HamEvo(gen, "t")))
    (op,) = conv.circuit.native.program.ops
    assert (op.tsym, op.steps, op.param) == ("t", 500, "duration")
    assert sorted(tm.expr for tm in op.generator.terms) == [f"sin({FY}*t)|t", "t**2*x|t,x"]
    assert bk.config._use_gate_params is False
    emb = conv.embedding_fn(conv.params, {"x": torch.tensor(2.0), "duration": torch.tensor(1.0)})
    out = oracle.run(conv.circuit.native.program, emb)[0].numpy()
    want = ode_solution(lambda s: OMEGA * (np.sin(FY * s) * np.kron(Xm, I2) + 2.0 * s**2 * np.kron(I2, Ym)), 4, 1.0)
    assert np.abs(out - want).max() < 5e-6


@pytest.mark.parametrize("duration", [0.5, 1.0])
def test_time_dependent_hamevo_through_quantum_model(oracle_device, duration: float) -> None:
    """Mirror of the reference's tests/backends/pyq/test_time_dependent_generator.py:25-62 with the ODE solution in
    place of qutip (not installed here) and the reference's own acceptance (MIDDLE_ACCEPTANCE = 6e-2 → we ask 1e-5)."""
    import sympy
    from qadence.parameters import Parameter, TimeParameter

    from tests.test_time_dependent import FY, OMEGA, X as Xm, Y as Ym, I2, ode_solution

    t, x = TimeParameter("t"), Parameter("x", trainable=False)
    hamevo = qadence.HamEvo(OMEGA * (sympy.sin(FY * t) * qadence.X(0) + x * (t**2) * qadence.Y(1)), "t")
    values = {"x": torch.tensor(2.0), "duration": torch.tensor(duration)}
    model = QuantumModel(QuantumCircuit(2, hamevo), backend="b200sv", diff_mode=DiffMode.AD, configuration={"n_steps_hevo": 500})
    state = model.run(values=values)
    want = ode_solution(lambda s: OMEGA * (np.sin(FY * s) * np.kron(Xm, I2) + 2.0 * s**2 * np.kron(I2, Ym)), 4, duration)
    assert np.abs(state[0].numpy() - want).max() < 1e-5
    state1 = qadence.run(hamevo, values=values, backend="b200sv", configuration={"n_steps_hevo": 500})
    assert torch.allclose(state, state1)


def _noise_ops():
    from qadence.blocks import MatrixBlock

    return [qadence.I(0) * qadence.I(1), qadence.X(0),
            MatrixBlock(block_to_tensor(qadence.X(0), qubit_support=(0, 1), use_full_support=True), qubit_support=(0, 1))]


@pytest.mark.parametrize("duration", [0.5, 1.0])
@pytest.mark.parametrize("which", [0, 1, 2])
def test_noisy_time_dependent_hamevo_through_quantum_model(oracle_device, duration: float, which: int) -> None:
    """Mirror of the reference's tests/backends/pyq/test_time_dependent_generator.py:68-126 (`HamEvo(...,
    noise_operators=[L])`, a density matrix comes back) with the integrated master equation in place of qutip `mesolve`
    (not installed here); the reference's acceptance is MIDDLE_ACCEPTANCE = 6e-2, we ask 1e-5."""
    import sympy
    from qadence.parameters import Parameter, TimeParameter

    from tests.test_lindblad import integrate_master_equation
    from tests.test_time_dependent import FY, OMEGA, X as Xm, Y as Ym, I2

    noise_op = _noise_ops()[which]
    t, x = TimeParameter("t"), Parameter("x", trainable=False)
    hamevo = qadence.HamEvo(OMEGA * (sympy.sin(FY * t) * qadence.X(0) + x * (t**2) * qadence.Y(1)), "t", noise_operators=[noise_op])
    values = {"x": torch.tensor(2.0), "duration": torch.tensor(duration)}
    model = QuantumModel(QuantumCircuit(2, hamevo), backend="b200sv", diff_mode=DiffMode.AD, configuration={"n_steps_hevo": 500})
    rho = model.run(values=values)
    assert type(rho).__name__ == "DensityMatrix" and rho.shape == (1, 4, 4)
    L = block_to_tensor(noise_op, qubit_support=(0, 1), use_full_support=True).squeeze(0).numpy()
    rho0 = np.zeros((4, 4), dtype=np.complex128)
    rho0[0, 0] = 1
    want = integrate_master_equation(lambda s: OMEGA * (np.sin(FY * s) * np.kron(Xm, I2) + 2.0 * s**2 * np.kron(I2, Ym)), [L], rho0, duration)
    assert np.abs(rho[0].numpy() - want).max() < 1e-5
    rho1 = qadence.run(hamevo, values=values, backend="b200sv", configuration={"n_steps_hevo": 500})
    assert torch.allclose(rho, rho1)


def test_noisy_constant_hamevo_conversion() -> None:
    """Constant generator + jump operators: exact semigroup, jump operators as dense matrices on the HamEvo support."""
    gen = 0.7 * qadence.X(0) + 0.4 * qadence.Z(0) * qadence.Z(2)
    bk = backend_factory("b200sv", diff_mode=None)
    conv = bk.convert(QuantumCircuit(3, qadence.HamEvo(gen, 1.3, noise_operators=[qadence.X(2), 0.5 * qadence.Z(0)])))
    (op,) = conv.circuit.native.program.ops
    assert op.targets == (0, 2) and len(op.jumps) == 2 and op.jumps[0].shape == (4, 4)
    assert np.allclose(op.jumps[0], np.kron(np.eye(2), [[0, 1], [1, 0]])) and np.allclose(op.jumps[1], 0.5 * np.kron(np.diag([1, -1]), np.eye(2)))
    assert conv.circuit.native.is_noisy


# ------------------------------------------------------------------------------------------ device-side embedding
@pytest.fixture(autouse=True)
def oracle_embedding_device(request: pytest.FixtureRequest, monkeypatch: pytest.MonkeyPatch) -> None:
    """Run the compiled embedding through its CPU restatement (test double for the CUDA kernels)."""
    from oracle import embedding as oemb
    from qadence_b200 import embedding as emb

    if request.node.name == "test_no_cpu_fallback":
        return

    monkeypatch.setattr(emb, "evaluate", oemb.evaluate)
    monkeypatch.setattr(emb, "_resolve_device", lambda device: torch.device("cpu"))


def _embedding_circuits():
    from qadence import BasisSet, ReuploadScaling, feature_map
    from qadence.parameters import FeatureParameter, VariationalParameter

    n = 3
    xs, w = FeatureParameter("x"), VariationalParameter("w")
    return [
        chain(feature_map(n, param="x", fm_type=BasisSet.CHEBYSHEV, reupload_scaling=ReuploadScaling.TOWER), hea(n, 2)),
        chain(feature_map(n, param="x", fm_type=BasisSet.FOURIER), feature_map(n, param="y", fm_type=BasisSet.CHEBYSHEV), hea(n, 1)),
        chain(qadence.RX(0, 2 * w * xs), qadence.RY(1, w), qadence.CRZ(0, 2, sympy_expr_sin(xs) + w**2), qadence.RX(2, 0.3) * (w + 1.5)),
    ]


def sympy_expr_sin(p):
    import sympy

    return sympy.sin(p)


@pytest.mark.parametrize("idx", [0, 1, 2])
@pytest.mark.parametrize("gate_level", [True, False])
def test_device_embedding_equals_reference_embedding(oracle_embedding_device, idx: int, gate_level: bool) -> None:
    """Same keys the program needs, same values, same gradients as qadence/blocks/embedding.py:118-167 — in
    expression-level (AD) and gate-level (ADJOINT) naming."""
    from qadence_b200.embedding import ParamTable

    block = _embedding_circuits()[idx]
    circ = QuantumCircuit(3, block)
    obs = [qadence.total_magnetization(3), 0.5 * qadence.VariationalParameter("ow") * qadence.Z(1)]
    on = backend_factory("b200sv", diff_mode=DiffMode.ADJOINT)
    off = backend_factory("b200sv", diff_mode=DiffMode.ADJOINT, configuration={"device_embedding": False})
    on.backend.config._use_gate_params = off.backend.config._use_gate_params = gate_level
    conv, ref = on.convert(circ, obs), off.convert(circ, obs)
    assert list(conv.params) == list(ref.params)
    rng = np.random.default_rng(idx)
    inputs = {"x": torch.tensor(rng.uniform(0.1, 0.9, size=5), requires_grad=True), "y": torch.tensor(rng.uniform(0.1, 0.9, size=5))}
    pa = {k: v.detach().clone().requires_grad_(v.requires_grad) for k, v in conv.params.items()}
    pb = {k: pa[k].detach().clone().requires_grad_(v.requires_grad) for k, v in ref.params.items()}  # (qadence re-draws observable parameters per convert)
    xa = {k: v.detach().clone().requires_grad_(v.requires_grad) for k, v in inputs.items()}
    xb = {k: v.detach().clone().requires_grad_(v.requires_grad) for k, v in inputs.items()}
    got, want = conv.embedding_fn(pa, xa), ref.embedding_fn(pb, xb)
    assert isinstance(got, ParamTable) and not isinstance(want, ParamTable)
    needed = list(conv.circuit.native.program.param_names) + [nm for o in conv.observable for nm in o.native.compiled.names]
    needed_ref = list(ref.circuit.native.program.param_names) + [nm for o in ref.observable for nm in o.native.compiled.names]
    assert needed and set(needed) <= set(got.names) and len(needed) == len(needed_ref)  # gate-level names are fresh uuids per convert
    loss_a = loss_b = 0.0
    for k, (nm, nm_ref) in enumerate(zip(needed, needed_ref)):
        a, b = got[nm], want[nm_ref]
        assert torch.allclose(a.expand(5) if a.numel() == 1 else a, (b.expand(5) if b.numel() == 1 else b).to(torch.float64), atol=1e-13, rtol=0), nm
        assert a.shape == b.reshape(-1).shape, (nm, a.shape, b.shape)  # [1] for batch-independent rows, like the reference
        loss_a = loss_a + (k + 1) * (a.expand(5) if a.numel() == 1 else a).sum()
        loss_b = loss_b + (k + 1) * (b.expand(5) if b.numel() == 1 else b).sum()
    loss_a.backward()
    loss_b.backward()
    for k in pa:
        if pa[k].requires_grad:
            ga = pa[k].grad if pa[k].grad is not None else torch.zeros_like(pa[k])
            gb = pb[k].grad if pb[k].grad is not None else torch.zeros_like(pb[k])
            assert torch.allclose(ga, gb, atol=1e-11), k
    assert torch.allclose(xa["x"].grad, xb["x"].grad, atol=1e-11)
    if not gate_level:  # expression-level dicts also pass the raw inputs and parameters through (embedding.py:155-166)
        assert torch.equal(got["x"], xa["x"]) and set(want) <= set(got.keys())


def test_device_embedding_through_quantum_model(oracle_device, oracle_embedding_device) -> None:
    from qadence import BasisSet, feature_map

    n = 3
    circ = QuantumCircuit(n, chain(feature_map(n, param="x", fm_type=BasisSet.CHEBYSHEV), hea(n, 2)))
    obs = qadence.total_magnetization(n)
    torch.manual_seed(0)
    m_on = QuantumModel(circ, obs, backend="b200sv", diff_mode=DiffMode.ADJOINT)
    m_off = QuantumModel(circ, obs, backend="b200sv", diff_mode=DiffMode.ADJOINT, configuration={"device_embedding": False})
    m_off.load_state_dict(m_on.state_dict())
    x = torch.linspace(0.1, 0.9, 4, dtype=torch.float64)
    e_on, e_off = m_on.expectation({"x": x}), m_off.expectation({"x": x})
    assert torch.allclose(e_on, e_off, atol=1e-12)
    e_on.sum().backward()
    e_off.sum().backward()
    n_checked = 0
    for (ka, a), (kb, b) in zip(m_on.named_parameters(), m_off.named_parameters()):
        if a.requires_grad:
            assert ka == kb and torch.allclose(a.grad, b.grad, atol=1e-10), ka
            n_checked += 1
    assert n_checked == 18
    assert torch.allclose(m_on.run({"x": x}), m_off.run({"x": x}), atol=1e-12)


# ------------------------------------------------------------------------------------------ shot-based measurements
@pytest.mark.parametrize("mode", [DiffMode.AD, DiffMode.GPSR])
def test_tomography_measurement_samples_from_this_backend(oracle_device, mode) -> None:
    """`measurement=Measurements("tomography")`: qadence's estimator (measurements/tomography.py:17-70) rotates the circuit
    into each Pauli basis and calls `backend.circuit` + `backend.sample` of THIS backend; the estimate converges to the
    exact expectation and, in GPSR mode, stays differentiable through the shift rules."""
    from qadence import Measurements

    circ = QuantumCircuit(2, chain(RX(0, "x"), RY(1, "t0"), CNOT(0, 1), RX(1, "t1")))
    obs = qadence.add(Z(0), 0.5 * kron(X(0), Z(1)))
    torch.manual_seed(3)
    shots = QuantumModel(circ, obs, backend="b200sv", diff_mode=mode, measurement=Measurements(protocol="tomography", options={"n_shots": 200000}))
    exact = QuantumModel(circ, obs, backend="b200sv", diff_mode=mode)
    exact.load_state_dict(shots.state_dict())
    v = {"x": torch.tensor([0.3, 0.9])}
    e_s, e_x = shots.expectation(v), exact.expectation(v)
    assert e_s.shape == e_x.shape == (2, 1)
    assert (e_s.detach() - e_x.detach()).abs().max() < 0.02  # ~ 4 sigma of two 200k-shot estimates
    if mode == DiffMode.GPSR:
        e_s.sum().backward()
        e_x.sum().backward()
        n_checked = 0
        for (ka, a), (kb, b) in zip(shots.named_parameters(), exact.named_parameters()):
            if a.requires_grad and b.grad is not None:
                assert a.grad is not None and (a.grad - b.grad).abs().max() < 0.03, ka
                n_checked += 1
        assert n_checked == 2


def test_classical_shadow_measurement(oracle_device) -> None:
    """`Measurements("shadow")` (measurements/shadow.py:124-175): one-shot samples of randomly rotated copies of the circuit,
    all converted and sampled by this backend."""
    from qadence import Measurements

    np.random.seed(0)
    circ = QuantumCircuit(2, chain(RX(0, "x"), RY(1, 0.4), CNOT(0, 1)))
    obs = qadence.add(Z(0), 0.5 * kron(X(0), X(1)))
    opts = {"accuracy": 0.3, "confidence": 0.1}
    shots = QuantumModel(circ, obs, backend="b200sv", diff_mode=DiffMode.AD, measurement=Measurements(protocol="shadow", options=opts))
    exact = QuantumModel(circ, obs, backend="b200sv", diff_mode=DiffMode.AD)
    v = {"x": torch.tensor([0.7])}
    assert (shots.expectation(v).detach() - exact.expectation(v).detach()).abs().max() < 0.45


@pytest.mark.parametrize("tool", ["cpu_fuzz_convert.py", "cpu_fuzz_observable.py"])
def test_randomised_converter_fuzz(tool: str) -> None:
    """Short runs of the converter campaigns (≈ 16 000 random blocks and 6 000 random observables were run once, clean)."""
    import subprocess
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parents[1]
    out = subprocess.run([sys.executable, str(root / "tools" / tool), "5", "6"], capture_output=True, text=True, timeout=300, cwd=root)
    last = [ln for ln in out.stdout.splitlines() if ln.strip()][-1:] or [""]
    assert out.returncode == 0 and last[0].startswith("ok:"), out.stdout[-1500:] + out.stderr[-1500:]


def test_pyqtorch_configuration_dicts_are_accepted(oracle_device) -> None:
    """Every option name of `backends/pyqtorch/config.py:38-90` (and `use_sparse_observable`, backend.py:43) is accepted, so
    user configuration dicts keep working; options that would change results raise instead of being ignored."""
    cfg = {"algo_hevo": "exp", "ode_solver": "dp5_se", "n_steps_hevo": 50, "use_gradient_checkpointing": True,
           "use_single_qubit_composition": True, "loop_expectation": True, "use_sparse_observable": True, "n_eqs": None,
           "shift_prefac": 0.5, "gap_step": 1.0, "lb": None, "ub": None, "dropout_probability": 0.0, "dropout_mode": "rotational"}
    model = QuantumModel(QuantumCircuit(2, chain(RX(0, "x"), CNOT(0, 1))), total_magnetization(2), backend="b200sv", configuration=cfg)
    assert model.expectation({"x": torch.tensor([0.2, 0.4])}).shape == (2, 1)


@pytest.mark.parametrize("mode", ["rotational_dropout", "entangling_dropout", "canonical_fwd_dropout"])
def test_quantum_dropout(oracle_device, mode: str) -> None:
    """`dropout_probability > 0` (pyqtorch/backend.py:127-141): in training mode every call executes a random sub-circuit,
    in eval mode the full circuit; a dropped trainable rotation gets no gradient in that pass."""
    from qadence_b200.qadence_backend.backend import dropout_ops

    n = 3
    circ = QuantumCircuit(n, chain(feature_map(n, param="x"), hea(n, 2)))
    torch.manual_seed(0)
    model = QuantumModel(circ, total_magnetization(n), backend="b200sv", diff_mode=DiffMode.ADJOINT,
                         configuration={"dropout_probability": 0.5, "dropout_mode": mode})
    plain = QuantumModel(circ, total_magnetization(n), backend="b200sv", diff_mode=DiffMode.ADJOINT)
    plain.load_state_dict(model.state_dict())
    assert model._circuit.native.dropout_prob == 0.5
    x = {"x": torch.tensor([0.3, 0.8])}
    model.eval()
    assert torch.allclose(model.expectation(x), plain.expectation(x), atol=1e-12)  # eval mode: the full circuit
    model.train()
    outs = {tuple(model.expectation(x).detach().flatten().round(decimals=9).tolist()) for _ in range(12)}
    assert len(outs) > 1  # different sub-circuits in different passes
    # the mask: only the gate classes of the mode are ever dropped, and only trainable rotations
    prog = model._circuit.native.program
    vals = model.embedding_fn(model._params, x)
    seen: set[int] = set()
    for _ in range(20):
        seen |= dropout_ops(prog, 0.5, mode, vals)
    for i in seen:
        op = prog.ops[i]
        if mode == "entangling_dropout":
            assert op.controls
        elif mode == "rotational_dropout":
            assert not op.controls and isinstance(op.param, str)
    assert seen
    e = model.expectation(x)
    e.sum().backward()
    assert any(p.grad is not None for p in model.parameters())


@pytest.mark.parametrize("dtype", [torch.complex64, torch.complex128])
def test_state_precision_comes_from_the_configuration(oracle_device, dtype) -> None:
    """Counterpart of tests/backends/pyq/test_quantum_pyq.py:875-903.  `QuantumModel.to(dtype)` only retargets pyqtorch
    circuits (model.py:651-655 checks `isinstance(native, pyqtorch.QuantumCircuit)`), so for this backend — as for any
    non-pyqtorch backend — the precision is a configuration option; complex-typed inputs are accepted like there."""
    circ = QuantumCircuit(3, chain(feature_map(3, param="x"), hea(3, 1)))
    name = "complex64" if dtype == torch.complex64 else "complex128"
    qm = QuantumModel(circ, total_magnetization(3), backend="b200sv", configuration={"dtype": name})
    assert qm._circuit.native.dtype == dtype and all(o.native.dtype == dtype for o in qm._observable)
    assert qm.to(dtype=dtype) is qm  # accepted, a no-op here
    inputs = {"x": torch.tensor([0.1, 0.5]).to(dtype=dtype)}
    state = qadence.uniform_state(3).to(dtype=dtype)
    wf = qm.run(inputs, state=state)
    assert wf.dtype == dtype and wf.shape == (2, 8)
    ev = qm.expectation(inputs, state=state)
    assert ev.dtype == (torch.float64 if dtype == torch.cdouble else torch.float32)


def test_default_and_custom_configuration() -> None:
    """tests/backends/test_backends.py:266-300 for this backend: default configuration object, `change_config` with a known
    option, ValueError for an unknown one."""
    from qadence.backend import BackendConfiguration

    bk = backend_factory("b200sv", diff_mode=None)
    conf = bk.default_configuration()
    assert isinstance(conf, BackendConfiguration) and isinstance(conf.available_options(), str) and "n_steps_hevo" in conf.available_options()
    model = QuantumModel(QuantumCircuit(1, RX(0, np.pi)), backend="b200sv", diff_mode=DiffMode.GPSR)
    old = model.backend.backend.config.use_sparse_observable
    model.change_config({"use_sparse_observable": not old})
    assert model.backend.backend.config.use_sparse_observable == (not old)
    with pytest.raises(ValueError):
        model.change_config({"sparse_observable": True})


def test_custom_transpilation_passes() -> None:
    """tests/backends/test_backends.py:316-337: user passes replace the default ones; an empty list disables them."""
    from qadence.backends.api import config_factory
    from qadence.transpile import flatten

    circuit = QuantumCircuit(1, chain(chain(chain(RX(0, np.pi / 2))), kron(kron(RX(0, np.pi / 2)))))
    config = config_factory("b200sv", {})
    config.transpilation_passes = [flatten]
    conv = backend_factory("b200sv", configuration=config).convert(circuit)
    config = config_factory("b200sv", {})
    config.transpilation_passes = []
    conv_none = backend_factory("b200sv", configuration=config).convert(circuit)
    assert conv.circuit.original == conv_none.circuit.original
    assert conv.circuit.abstract != conv_none.circuit.abstract
    assert len(conv.circuit.native.program.ops) == len(conv_none.circuit.native.program.ops) == 2


# ------------------------------------------------------------------------------------------ digital / readout noise
def _dense_noisy_density(block, n: int, noise_after: dict, values: dict | None = None) -> np.ndarray:
    """ρ of a chain of primitive blocks from the reference's dense matrices (block_to_tensor) and explicit Kraus sums."""
    from qadence_b200 import density

    rho = np.zeros((2**n, 2**n), dtype=np.complex128)
    rho[0, 0] = 1.0
    for k, b in enumerate(block.blocks if hasattr(block, "blocks") else [block]):
        u = block_to_tensor(b, values or {}, qubit_support=tuple(range(n)))[0].numpy()
        rho = u @ rho @ u.conj().T
        for name, probs in noise_after.get(k, ()):
            q = b.qubit_support[-1]
            ks = [np.kron(np.kron(np.eye(2**q), kk), np.eye(2 ** (n - q - 1))) for kk in density.kraus_operators(name, probs)]
            rho = sum(kk @ rho @ kk.conj().T for kk in ks)
    return rho


@pytest.mark.parametrize("noisy_config", ["BITFLIP", ["BITFLIP", "PHASEFLIP"]])
def test_run_digital_noise_like_the_reference_test(oracle_device: None, noisy_config) -> None:
    """Mirror of /root/reference/tests/qadence/test_noise/test_digital_noise.py:74-98 on this backend, plus the numbers."""
    from qadence import NoiseHandler, NoiseProtocol, hamiltonian_factory, set_noise

    protos = [getattr(NoiseProtocol.DIGITAL, p) for p in (noisy_config if isinstance(noisy_config, list) else [noisy_config])]
    block = kron(H(0), Z(1))
    circuit = QuantumCircuit(2, block)
    observable = hamiltonian_factory(circuit.n_qubits, detuning=Z)
    noise = NoiseHandler(protos if len(protos) > 1 else protos[0], {"error_probability": 0.1})
    model = QuantumModel(circuit=circuit, observable=observable, backend="b200sv", diff_mode=DiffMode.GPSR)
    noiseless_output = model.run()
    assert noiseless_output.shape == (1, 4)
    set_noise(circuit, noise)
    noisy_model = QuantumModel(circuit=circuit, observable=observable, backend="b200sv", diff_mode=DiffMode.GPSR)
    noisy_output = noisy_model.run()
    assert type(noisy_output).__name__ == "DensityMatrix" and noisy_output.shape == (1, 4, 4)
    pure = torch.einsum("bi,bj->bij", noiseless_output, noiseless_output.conj())
    assert not torch.allclose(pure, noisy_output.as_subclass(torch.Tensor))
    backend = backend_factory(backend="b200sv", diff_mode=DiffMode.GPSR)
    (conv_circ, _, embed, params) = backend.convert(circuit)
    native_output = backend.run(conv_circ, embed(params, {}))
    assert torch.allclose(noisy_output.as_subclass(torch.Tensor), native_output.as_subclass(torch.Tensor))
    each = tuple((str(p.value), (0.1,)) for p in protos)
    want = _dense_noisy_density(block, 2, {0: each, 1: each})
    assert np.abs(noisy_output[0].as_subclass(torch.Tensor).numpy() - want).max() < 1e-12
    assert abs(np.trace(want) - 1) < 1e-12


@pytest.mark.parametrize("noisy_config", ["BITFLIP", ["BITFLIP", "PHASEFLIP"]])
def test_expectation_digital_noise_like_the_reference_test(oracle_device: None, noisy_config) -> None:
    """Mirror of test_digital_noise.py:101-127: `noise=` at expectation time is sticky on the converted circuit."""
    from qadence import NoiseHandler, NoiseProtocol, hamiltonian_factory

    protos = [getattr(NoiseProtocol.DIGITAL, p) for p in (noisy_config if isinstance(noisy_config, list) else [noisy_config])]
    block = kron(H(0), Z(1))
    circuit = QuantumCircuit(2, block)
    observable = hamiltonian_factory(circuit.n_qubits, detuning=Z)
    noise = NoiseHandler(protos if len(protos) > 1 else protos[0], {"error_probability": 0.1})
    backend = backend_factory(backend="b200sv", diff_mode=DiffMode.GPSR)
    model = QuantumModel(circuit=circuit, observable=observable, backend="b200sv", diff_mode=DiffMode.GPSR)
    noiseless_expectation = model.expectation(values={})
    (conv_circ, conv_obs, embed, params) = backend.convert(circuit, observable)
    native_noisy_expectation = backend.expectation(conv_circ, conv_obs, embed(params, {}), noise=noise)
    assert not torch.allclose(noiseless_expectation, native_noisy_expectation)
    noisy_model = QuantumModel(circuit=circuit, observable=observable, noise=noise, backend="b200sv", diff_mode=DiffMode.GPSR)
    assert torch.allclose(noisy_model.expectation(values={}), native_noisy_expectation)
    (conv_circ, conv_obs, embed, params) = backend.convert(circuit, observable)
    assert torch.allclose(backend.expectation(conv_circ, conv_obs, embed(params, {})), native_noisy_expectation)
    # numbers: Tr(O rho) with the dense Kraus evolution
    each = tuple((str(p.value), (0.1,)) for p in protos)
    rho = _dense_noisy_density(block, 2, {0: each, 1: each})
    o = block_to_tensor(observable, {}, qubit_support=(0, 1))[0].numpy()
    assert abs(float(native_noisy_expectation) - np.trace(o @ rho).real) < 1e-12


@pytest.mark.parametrize("protocol,probs", [("BITFLIP", 0.2), ("PHASEFLIP", 0.15), ("DEPOLARIZING", 0.3), ("PAULI_CHANNEL", (0.1, 0.2, 0.05)),
                                            ("AMPLITUDE_DAMPING", 0.4), ("PHASE_DAMPING", 0.35), ("GENERALIZED_AMPLITUDE_DAMPING", (0.6, 0.3))])
def test_every_digital_protocol_on_a_parametric_circuit(oracle_device: None, protocol: str, probs) -> None:
    from qadence import NoiseHandler, NoiseProtocol

    n = 3
    block = chain(RX(0, "a"), RY(1, "b"), CNOT(0, 1), RZ(2, "a"), CRX(1, 2, "b"), H(0), CPHASE(2, 0, 0.7))
    noise = NoiseHandler(getattr(NoiseProtocol.DIGITAL, protocol), {"error_probability": probs})
    obs = [add(Z(0), Z(1), Z(2)), kron(X(0), Y(2))]
    from qadence_b200.qadence_backend.config import Configuration

    backend = backend_factory(backend="b200sv", diff_mode=None, configuration=Configuration(noise=noise))
    conv = backend.convert(QuantumCircuit(n, block), obs)
    vals = {"a": torch.tensor([0.3, 1.1]), "b": torch.tensor([-0.8, 2.0])}
    rho = backend.run(conv.circuit, conv.embedding_fn(conv.params, vals))
    ex = backend.expectation(conv.circuit, conv.observable, conv.embedding_fn(conv.params, vals))
    assert rho.shape == (2, 8, 8) and ex.shape == (2, 2)
    p = tuple(probs) if isinstance(probs, tuple) else (probs,)
    name = str(getattr(NoiseProtocol.DIGITAL, protocol).value)
    for b in range(2):
        v = {k: t[b:b + 1] for k, t in vals.items()}
        want = _dense_noisy_density(block, n, {k: ((name, p),) for k in range(7)}, v)
        assert np.abs(rho[b].as_subclass(torch.Tensor).numpy() - want).max() < 1e-12
        for j, o in enumerate(obs):
            om = block_to_tensor(o, {}, qubit_support=tuple(range(n)))[0].numpy()
            assert abs(float(ex[b, j]) - np.trace(om @ want).real) < 1e-12


def test_gpsr_gradient_of_a_noisy_expectation(oracle_device: None) -> None:
    """The parameter-shift rules hold for noisy expectations: dE/dθ by GPSR == finite difference of Tr(O ρ(θ))."""
    from qadence import NoiseHandler, NoiseProtocol

    block = chain(RX(0, "th"), CNOT(0, 1), RY(1, 0.4))
    noise = NoiseHandler(NoiseProtocol.DIGITAL.DEPOLARIZING, {"error_probability": 0.15})
    model = QuantumModel(QuantumCircuit(2, block), observable=add(Z(0), Z(1)), backend="b200sv", diff_mode=DiffMode.GPSR, noise=noise)
    th = [p_ for p_ in model.parameters() if p_.requires_grad][0]
    e = model.expectation({})
    (g,) = torch.autograd.grad(e.sum(), th)

    def value(t: float) -> float:
        each = (("Depolarizing", (0.15,)),)
        rho = _dense_noisy_density(chain(RX(0, t), CNOT(0, 1), RY(1, 0.4)), 2, {0: each, 1: each, 2: each})
        o = block_to_tensor(add(Z(0), Z(1)), {}, qubit_support=(0, 1))[0].numpy()
        return float(np.trace(o @ rho).real)

    t0 = float(th.detach().reshape(-1)[0])
    assert abs(float(e) - value(t0)) < 1e-12
    assert abs(float(g.reshape(-1)[0]) - (value(t0 + 1e-6) - value(t0 - 1e-6)) / 2e-6) < 1e-8


def test_readout_noise_changes_the_sample_distribution(oracle_device: None) -> None:
    """Statistical contract of /root/reference/tests/qadence/test_noise/test_readout.py:67-110."""
    from qadence import NoiseHandler, NoiseProtocol

    circuit = QuantumCircuit(2, kron(X(0), X(1)))
    model = QuantumModel(circuit, backend="b200sv", diff_mode=DiffMode.GPSR)
    clean = model.sample(n_shots=400)[0]
    assert clean == Counter({"11": 400})
    noise = NoiseHandler(NoiseProtocol.READOUT.INDEPENDENT, {"error_probability": 0.25, "seed": 0})
    noisy = model.sample(n_shots=4000, noise=noise)[0]
    assert sum(noisy.values()) == 4000 and set(noisy) == {"00", "01", "10", "11"}
    assert abs(noisy["11"] / 4000 - 0.75**2) < 0.04 and abs(noisy["00"] / 4000 - 0.25**2) < 0.03
    cm = torch.zeros(4, 4)
    cm[0, 3] = 1.0  # every "11" is read as "00"
    cm[1, 1] = cm[2, 2] = cm[3, 0] = 1.0
    corr = NoiseHandler(NoiseProtocol.READOUT.CORRELATED, {"confusion_matrix": cm, "seed": 0})
    model2 = QuantumModel(QuantumCircuit(2, kron(X(0), X(1))), backend="b200sv", diff_mode=DiffMode.GPSR)
    assert model2.sample(n_shots=100, noise=corr)[0] == Counter({"00": 100})


def test_density_matrix_initial_state_and_little_endian(oracle_device: None) -> None:
    from qadence_b200.density import DensityMatrix

    block = chain(RX(0, 0.3), CNOT(0, 1), RY(2, 1.2))
    backend = backend_factory(backend="b200sv", diff_mode=None)
    conv = backend.convert(QuantumCircuit(3, block), [Z(0) + Z(2)])
    psi0 = torch.randn(2, 8, dtype=torch.cdouble)
    psi0 = psi0 / psi0.norm(dim=1, keepdim=True)
    rho0 = torch.einsum("bi,bj->bij", psi0, psi0.conj()).as_subclass(DensityMatrix)
    psi = backend.run(conv.circuit, {}, state=psi0)
    rho = backend.run(conv.circuit, {}, state=rho0)
    assert type(rho).__name__ == "DensityMatrix"
    assert torch.allclose(rho.as_subclass(torch.Tensor), torch.einsum("bi,bj->bij", psi, psi.conj()), atol=1e-12)
    e_psi = backend.expectation(conv.circuit, conv.observable, {}, state=psi0)
    e_rho = backend.expectation(conv.circuit, conv.observable, {}, state=rho0)
    assert torch.allclose(e_psi, e_rho, atol=1e-12)
    little = backend.run(conv.circuit, {}, state=rho0, endianness=Endianness.LITTLE).as_subclass(torch.Tensor)
    psi_l = backend.run(conv.circuit, {}, state=psi0, endianness=Endianness.LITTLE)
    assert torch.allclose(little, torch.einsum("bi,bj->bij", psi_l, psi_l.conj()), atol=1e-12)
    with pytest.raises(ValueError):
        backend.run(conv.circuit, {}, state=torch.zeros(1, 4, 4, dtype=torch.cdouble).as_subclass(DensityMatrix))


@pytest.mark.parametrize("mode", [DiffMode.AD, DiffMode.ADJOINT])
def test_noisy_expectation_is_differentiable_in_ad_and_adjoint_mode(oracle_device: None, mode) -> None:
    """Tr(O rho) is a linear functional of the vectorised density matrix: autograd through the differentiable `run` of the
    vectorised program == GPSR == finite differences of the dense Kraus evolution."""
    from qadence import NoiseHandler, NoiseProtocol

    block = chain(RX(0, "th"), CNOT(0, 1), RY(1, 0.4), CRZ(1, 0, "ph"))
    noise = NoiseHandler([NoiseProtocol.DIGITAL.DEPOLARIZING, NoiseProtocol.DIGITAL.AMPLITUDE_DAMPING], [{"error_probability": 0.15}, {"error_probability": 0.2}])
    obs = [add(Z(0), Z(1)), kron(X(0), Y(1))]
    model = QuantumModel(QuantumCircuit(2, block), observable=obs, backend="b200sv", diff_mode=mode, noise=noise)
    ref = QuantumModel(QuantumCircuit(2, block), observable=obs, backend="b200sv", diff_mode=DiffMode.GPSR, noise=noise)
    ref.load_state_dict(model.state_dict())
    e, e_ref = model.expectation({}), ref.expectation({})
    assert e.requires_grad and torch.allclose(e, e_ref, atol=1e-12)
    w = torch.tensor([[0.7, -1.3]])
    g = torch.autograd.grad((e * w).sum(), [p_ for p_ in model.parameters() if p_.requires_grad])
    g_ref = torch.autograd.grad((e_ref * w).sum(), [p_ for p_ in ref.parameters() if p_.requires_grad])
    assert len(g) == 2
    for a, b in zip(g, g_ref):
        assert torch.allclose(a, b, atol=1e-10), (a, b)
    vals = {k: float(v) for k, v in model.vparams.items()}

    def value(th: float, ph: float) -> float:
        each = (("Depolarizing", (0.15,)), ("AmplitudeDamping", (0.2,)))
        rho = _dense_noisy_density(chain(RX(0, th), CNOT(0, 1), RY(1, 0.4), CRZ(1, 0, ph)), 2, {k: each for k in range(4)})
        return float(sum(wt * np.trace(block_to_tensor(o, {}, qubit_support=(0, 1))[0].numpy() @ rho).real for wt, o in zip([0.7, -1.3], obs)))

    names = [k for k, p_ in model._params.items() if p_.requires_grad]
    for gi, nm in zip(g, names):
        hi = value(vals["th"] + (1e-6 if nm == "th" else 0), vals["ph"] + (1e-6 if nm == "ph" else 0))
        lo = value(vals["th"] - (1e-6 if nm == "th" else 0), vals["ph"] - (1e-6 if nm == "ph" else 0))
        assert abs(float(gi) - (hi - lo) / 2e-6) < 1e-7, nm


def test_algo_hevo_trotter_reaches_the_converted_program() -> None:
    """`Configuration(algo_hevo="trotter", n_steps_hevo=k)` marks Pauli-sum HamEvo ops for the Trotter propagator (Op.steps = k);
    the default and the reference's own AlgoHEvo names keep the exact one (steps = 0)."""
    from qadence_b200.program import OpKind

    block = HamEvo(add(0.7 * kron(Z(0), Z(1)), 0.3 * X(0), 0.2 * Y(1)), 0.9)
    for cfg, want in (({}, 0), ({"algo_hevo": "EXP"}, 0), ({"algo_hevo": "trotter", "n_steps_hevo": 25}, 25)):
        bk = backend_factory("b200sv", diff_mode=None, configuration=cfg)
        ops = bk.convert(QuantumCircuit(2, block)).circuit.native.program.ops
        evo = [o for o in ops if o.kind == OpKind.HAMEVO]
        assert len(evo) == 1 and evo[0].steps == want, (cfg, evo[0].steps)
