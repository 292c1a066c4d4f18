"""Edge cases the reference's own tests exercise (file:line in each docstring), through the CUDA path."""

from __future__ import annotations

import numpy as np
import pytest
import torch

from oracle import statevec as oracle
from qadence_b200 import engine
from qadence_b200.program import CONST_MATS, Op, OpKind, PauliSum, PauliTerm, Program, ProgramBuilder, hea_program, total_magnetization
from tests.test_plan_cpu import random_program

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rand_state(rng, B, n):
    s = rng.normal(size=(B, 2**n)) + 1j * rng.normal(size=(B, 2**n))
    s /= np.linalg.norm(s, axis=1, keepdims=True)
    return torch.tensor(s)


def test_empty_circuit_and_identity() -> None:
    """tests/backends/test_endianness.py:116 — QuantumCircuit(3) leaves |000>."""
    cp = engine.CompiledProgram(Program(3, []))
    out = engine.run(cp, {}, device=DEV).cpu()
    assert out.shape == (1, 8) and out[0, 0] == 1 and out.abs().sum() == 1
    cp = engine.CompiledProgram(ProgramBuilder(3).gate("I", 1).build())
    assert torch.equal(engine.run(cp, {}, device=DEV).cpu(), out)


def test_state_batch_broadcast_and_dtypes() -> None:
    """One initial state, batched parameters (backends/pyqtorch/backend.py:158-179); complex64 in → complex64 out,
    float32 expectations (tests/backends/pyq/test_quantum_pyq.py:873-901)."""
    rng = np.random.default_rng(0)
    n, B = 5, 7
    prog = hea_program(n, 1)
    vals = {nm: torch.tensor(rng.uniform(0, 6, size=B)) for nm in prog.param_names}
    psi0 = _rand_state(rng, 1, n)
    want = oracle.run(prog, vals, psi0)
    cp = engine.CompiledProgram(prog)
    got = engine.run(cp, vals, psi0, device=DEV)
    assert got.dtype == torch.complex128 and (got.cpu() - want).abs().max() < 1e-12
    got32 = engine.run(cp, vals, psi0.to(torch.complex64), dtype=torch.complex64, device=DEV)
    assert got32.dtype == torch.complex64 and (got32.cpu().to(torch.complex128) - want).abs().max() < 1e-5
    e32 = engine.expectation(cp, [engine.CompiledObservable(total_magnetization(n))], vals, psi0, dtype=torch.complex64, device=DEV)
    assert e32.dtype == torch.float32 and e32.shape == (B, 1)


@pytest.mark.parametrize("batch", [1, 2, 5, 9])
def test_scaled_feature_parameter_batching(batch: int) -> None:
    """tests/backends/pyq/test_quantum_pyq.py:684-696 — `x * X(0)`: wf(x) == x · wf(1)."""
    prog = ProgramBuilder(1).X(0).scale("x").build()
    cp = engine.CompiledProgram(prog)
    x = torch.rand(batch, dtype=torch.float64) + 0.1
    wf = engine.run(cp, {"x": x}, device=DEV).cpu()
    wf1 = engine.run(cp, {"x": torch.ones(batch, dtype=torch.float64)}, device=DEV).cpu()
    assert torch.allclose(wf, wf1 * x.unsqueeze(1).to(torch.complex128))


def test_circuit_followed_by_its_dagger_is_identity() -> None:
    """tests/backends/pyq/test_quantum_pyq.py:699-756 — block · block.dagger() == 1 for every gate family."""
    rng = np.random.default_rng(4)
    n = 6
    prog, names = random_program(n, 50, rng)
    ops = [op for op in prog.ops if op.kind != OpKind.SCALE and not (op.kind == OpKind.MAT1 and op.label in ("rand", "N", "P0", "P1"))]
    dag = []
    for op in reversed(ops):
        if op.kind in (OpKind.RX, OpKind.RY, OpKind.RZ, OpKind.PHASE):
            if isinstance(op.param, str):
                dag.append(Op(op.kind, op.targets, op.controls, param=op.param + "_neg", label=op.label))
            else:
                dag.append(Op(op.kind, op.targets, op.controls, param=-op.param, label=op.label))
        elif op.kind == OpKind.MAT1:
            dag.append(Op(op.kind, op.targets, op.controls, matrix=np.asarray(op.matrix).conj().T, label=op.label))
        else:
            dag.append(op)
    vals = {nm: torch.tensor(rng.uniform(0.1, 3, size=2)) for nm in names}
    vals.update({nm + "_neg": -v for nm, v in list(vals.items())})
    psi0 = _rand_state(rng, 2, n)
    out = engine.run(engine.CompiledProgram(Program(n, ops + dag)), vals, psi0, device=DEV).cpu()
    assert (out - psi0).abs().max() < 1e-12


def test_high_and_low_qubits_on_a_20_qubit_register() -> None:
    """Targets / controls on the most and least significant index bits at cfg2's register size."""
    n = 20
    rng = np.random.default_rng(8)
    b = ProgramBuilder(n)
    for q in (0, 1, 19, 18, 9, 10):
        b.RX(q, "a").RY(q, 0.37).H(q)
    b.CNOT(0, 19).CNOT(19, 0).CRZ(1, 18, "a").CPHASE(18, 1, 0.9).SWAP(0, 19).Toffoli((0, 19), 10).CZ(9, 0)
    prog = b.build()
    vals = {"a": torch.tensor([0.77])}
    want = oracle.run(prog, vals)
    got = engine.run(engine.CompiledProgram(prog), vals, device=DEV).cpu()
    assert (got - want).abs().max() < 1e-12


def test_add_block_in_circuit_and_multi_qubit_projector() -> None:
    """tests/backends/pyq/test_quantum_pyq.py:789-802 — kron(P0, I) + kron(P1, X) == CNOT; Projector on (0, 2)."""
    n = 3
    rng = np.random.default_rng(1)
    psi0 = _rand_state(rng, 2, n)
    cnot_add = ProgramBuilder(n).add([
        ProgramBuilder(n).gate("P0", 0).build(),
        ProgramBuilder(n).gate("P1", 0).X(1).build(),
    ]).build()
    got = engine.run(engine.CompiledProgram(cnot_add), {}, psi0, device=DEV).cpu()
    want = engine.run(engine.CompiledProgram(ProgramBuilder(n).CNOT(0, 1).build()), {}, psi0, device=DEV).cpu()
    assert (got - want).abs().max() < 1e-14
    m = np.zeros((4, 4), dtype=complex)
    m[0b10, 0b01] = 1.0  # |10><01| on qubits (0, 2)
    proj = ProgramBuilder(n).H(1).matrix(m, (0, 2)).build()
    assert (engine.run(engine.CompiledProgram(proj), {}, psi0, device=DEV).cpu() - oracle.run(proj, {}, psi0)).abs().max() < 1e-14


def test_generic_program_observable_and_parametric_observable_gradient() -> None:
    """AD mode accepts parametric observables (engines/differentiable_backend.py:136-148): dE/dw = <psi|P|psi>."""
    n, B = 4, 3
    rng = np.random.default_rng(2)
    prog = hea_program(n, 1)
    vals = {nm: torch.tensor(rng.uniform(0, 6, size=B)).requires_grad_(True) for nm in prog.param_names}
    w = torch.tensor([0.8], dtype=torch.float64, requires_grad=True)
    obs = PauliSum(n, [PauliTerm(1.0, ("w",), ((0, "Z"),)), PauliTerm(0.5, (), ((1, "X"), (2, "X"))), PauliTerm(0.25, ("w",), ((3, "Y"),))])
    cp = engine.CompiledProgram(prog)
    e = engine.expectation(cp, [engine.CompiledObservable(obs)], {**vals, "w": w}, device=DEV)
    e.sum().backward()
    det = {k: v.detach() for k, v in vals.items()}
    ref = oracle.expectation(prog, [obs], {**det, "w": w.detach()})
    assert (e.detach().cpu() - ref).abs().max() < 1e-12
    dz = oracle.expectation(prog, [PauliSum(n, [PauliTerm(1.0, (), ((0, "Z"),)), PauliTerm(0.25, (), ((3, "Y"),))])], det)
    assert abs(w.grad.item() - dz.sum().item()) < 1e-10
    # a non-Pauli observable given as a generic program: Re<psi| CNOT |psi>
    gobs = ProgramBuilder(n).CNOT(0, 1).build()
    e2 = engine.expectation(cp, [engine.CompiledObservable(gobs)], det, device=DEV).cpu()
    assert (e2 - oracle.expectation(prog, [gobs], det)).abs().max() < 1e-12


def test_hamevo_parametric_generator_coefficient() -> None:
    n, B = 4, 3
    rng = np.random.default_rng(3)
    gen = PauliSum(n, [PauliTerm(0.9, ("g",), ((0, "X"), (1, "X"))), PauliTerm(0.4, (), ((2, "Z"), (3, "Z"))), PauliTerm(1.0, ("g",), ((3, "Y"),))])
    prog = ProgramBuilder(n).H(0).hamevo(gen, "t").build()
    vals = {"g": torch.tensor(rng.uniform(0.5, 1.5, size=B)), "t": torch.tensor(rng.uniform(0.2, 2.0, size=B))}
    psi0 = _rand_state(rng, B, n)
    got = engine.run(engine.CompiledProgram(prog), vals, psi0, device=DEV).cpu()
    assert (got - oracle.run(prog, vals, psi0)).abs().max() < 1e-10


def test_sampling_statistics_and_total_counts() -> None:
    """tests/backends/pyq/test_quantum_pyq.py:607-617,139-153 — H|0>: 40..60 of 100; every Counter sums to n_shots."""
    cp = engine.CompiledProgram(ProgramBuilder(1).H(0).build())
    torch.manual_seed(0)
    (c,) = engine.sample(cp, n_shots=100, device=DEV)
    assert 35 <= c["0"] <= 65 and c["0"] + c["1"] == 100
    prog = hea_program(4, 1)
    vals = {nm: torch.rand(5, dtype=torch.float64) for nm in prog.param_names}
    cs = engine.sample(engine.CompiledProgram(prog), vals, n_shots=77, device=DEV)
    assert len(cs) == 5 and all(sum(c.values()) == 77 for c in cs)
    little = engine.sample(engine.CompiledProgram(ProgramBuilder(3).X(0).build()), n_shots=10, device=DEV, little_endian=True)
    assert little == [{"001": 10}]


def test_missing_parameter_and_bad_state_errors() -> None:
    cp = engine.CompiledProgram(ProgramBuilder(2).RX(0, "theta").build())
    with pytest.raises(KeyError):
        engine.run(cp, {}, device=DEV)
    with pytest.raises(ValueError):
        engine.run(cp, {"theta": torch.zeros(1)}, torch.zeros(1, 8, dtype=torch.complex128), device=DEV)
    with pytest.raises(ValueError):
        engine.run(cp, {"theta": torch.zeros(3)}, torch.zeros(2, 4, dtype=torch.complex128), device=DEV)


def test_observable_parameter_gradient_without_circuit_parameters() -> None:
    """dE/dw for a parametric observable when nothing in the circuit asks for a gradient (found by tools/gpu_fuzz_mixed.py:
    the no-adjoint branch used to drop it)."""
    n = 3
    prog = ProgramBuilder(n).H(0).CNOT(0, 1).RX(2, 0.4).build()
    obs = PauliSum(n, [PauliTerm(0.7, ("w",), ((0, "Z"), (1, "Z"))), PauliTerm(1.0, (), ((2, "Z"),)), PauliTerm(-0.3, ("w", "w"), ((2, "Y"),))])
    w = torch.tensor([0.9, -0.4], dtype=torch.float64, device=DEV, requires_grad=True)
    e = engine.expectation(engine.CompiledProgram(prog), [engine.CompiledObservable(obs)], {"w": w}, device=DEV)
    e.sum().backward()
    wh = w.detach().cpu().requires_grad_(True)
    ref = oracle.expectation(prog, [obs], {"w": wh})
    ref.sum().backward()
    assert (e.detach().cpu() - ref.detach()).abs().max() < 1e-12
    assert (w.grad.cpu() - wh.grad).abs().max() < 1e-12


@pytest.mark.parametrize("diagonal", [True, False])
def test_hamevo_generator_coefficient_gradients(diagonal: bool) -> None:
    """Adjoint gradients w.r.t. parametric coefficients of a Pauli-sum HamEvo generator (analog blocks with parametric
    omega / delta): constant integrand for commuting generators, Gauss-Legendre along the evolution otherwise."""
    n, B = 4, 3
    rng = np.random.default_rng(11)
    p = "Z" if diagonal else "X"
    gen = PauliSum(n, [PauliTerm(0.9, ("g",), ((0, "Z"), (1, "Z"))), PauliTerm(0.4, ("h",), ((2, p),)), PauliTerm(1.1, ("g", "h"), ((3, "Z"),)),
                       PauliTerm(-0.6, (), ((1, p), (2, "Z")))])
    prog = ProgramBuilder(n).H(0).RY(1, "a").hamevo(gen, "t").CNOT(0, 3).RX(2, "a").build()
    obs = PauliSum(n, [PauliTerm(1.0, (), ((0, "X"),)), PauliTerm(0.5, (), ((1, "Z"), (3, "Y")))])
    host = {"g": torch.tensor(rng.uniform(0.5, 1.5, size=B), requires_grad=True), "h": torch.tensor([0.8], dtype=torch.float64, requires_grad=True),
            "t": torch.tensor(rng.uniform(0.3, 2.5, size=B), requires_grad=True), "a": torch.tensor([0.6], dtype=torch.float64, requires_grad=True)}
    want = oracle.expectation(prog, [obs], host)
    want.sum().backward()
    devv = {k: v.detach().clone().to(DEV).requires_grad_(True) for k, v in host.items()}
    got = engine.expectation(engine.CompiledProgram(prog), [engine.CompiledObservable(obs)], devv, device=DEV)
    got.sum().backward()
    assert (got.detach().cpu() - want.detach()).abs().max() < 1e-10
    for k in host:
        assert (devv[k].grad.cpu() - host[k].grad).abs().max() < 1e-9, k
