"""Density-matrix / digital-noise path (qadence_b200/density.py) on CPU.

The vectorised program (2n qubits) is executed by the CPU oracle and compared with an INDEPENDENT dense evolution
ρ → U ρ U†, ρ → Σ K ρ K† built from full 2^n × 2^n matrices, plus closed-form known answers of the textbook channels.
GPU execution of the same transformed programs: tests/test_gpu_density.py.
"""

from __future__ import annotations

import numpy as np
import pytest
import torch

from oracle import statevec as oracle
from qadence_b200 import density
from qadence_b200.program import CONST_MATS, Op, OpKind, PauliSum, PauliTerm, Program, ProgramBuilder

PROTOCOLS = [("BitFlip", (0.13,)), ("PhaseFlip", (0.21,)), ("Depolarizing", (0.17,)), ("PauliChannel", (0.05, 0.11, 0.2)),
             ("AmplitudeDamping", (0.3,)), ("PhaseDamping", (0.25,)), ("GeneralizedAmplitudeDamping", (0.4, 0.35))]


def embed(m: np.ndarray, q: int, n: int) -> np.ndarray:
    return np.kron(np.kron(np.eye(2**q), m), np.eye(2 ** (n - q - 1)))


def op_unitary(op: Op, n: int, values: dict, b: int) -> np.ndarray:
    """Full matrix of one noise-free op for batch element b: the (pinned) oracle applied to the identity's columns."""
    vals = {k: (v[b:b + 1] if isinstance(v, torch.Tensor) and v.dim() >= 1 and v.shape[0] > 1 else v) for k, v in values.items()}
    clean = Op(op.kind, op.targets, op.controls, param=op.param, matrix=op.matrix, subs=op.subs, generator=op.generator,
               label=op.label, tsym=op.tsym, steps=op.steps)
    cols = oracle.run(Program(n, [clean]), vals, torch.eye(2**n, dtype=torch.complex128))
    return cols.numpy().T  # row k of `cols` = U e_k


def dense_density(prog: Program, values: dict, rho0: np.ndarray, b: int) -> np.ndarray:
    n = prog.n_qubits
    rho = rho0.copy()
    for op in prog.ops:
        u = op_unitary(op, n, values, b)
        rho = u @ rho @ u.conj().T
        for name, probs in op.noise:
            ks = [embed(k, op.targets[0], n) for k in density.kraus_operators(name, probs)]
            rho = sum(k @ rho @ k.conj().T for k in ks)
    return rho


def run_vectorised(prog: Program, values: dict, state=None) -> torch.Tensor:
    vprog, needs = density.vectorise(prog)
    n = prog.n_qubits
    v = oracle.run(vprog, density.extend_values(values, needs), density.vectorise_state(state, n))
    return v.reshape(v.shape[0], 2**n, 2**n)


def random_noisy_program(rng: np.random.Generator, n: int, n_ops: int, with_hamevo: bool = True) -> tuple[Program, dict]:
    b = ProgramBuilder(n)
    values: dict = {}
    B = 3

    def param():
        if rng.random() < 0.3:
            return float(rng.uniform(-3, 3))
        nm = f"p{len(values)}"
        values[nm] = torch.tensor(rng.uniform(-3, 3, size=B if rng.random() < 0.5 else 1))
        return nm

    for _ in range(n_ops):
        kind = rng.integers(0, 10 if with_hamevo else 8)
        q = int(rng.integers(0, n))
        others = [x for x in range(n) if x != q]
        noisy_ok = True
        if kind == 0:
            b.rot([OpKind.RX, OpKind.RY, OpKind.RZ, OpKind.PHASE][rng.integers(0, 4)], q, param())
        elif kind == 1:
            b.gate(["X", "Y", "Z", "H", "S", "T", "SDagger"][rng.integers(0, 7)], q)
        elif kind == 2 and others:
            b.gate(["X", "Z", "Y"][rng.integers(0, 3)], q, (int(rng.choice(others)),))
        elif kind == 3 and others:
            b.rot([OpKind.RX, OpKind.RY, OpKind.RZ, OpKind.PHASE][rng.integers(0, 4)], q, param(), (int(rng.choice(others)),))
        elif kind == 4 and others:
            b.SWAP(q, int(rng.choice(others)))
        elif kind == 5:
            b.scale(complex(rng.uniform(0.5, 1.5), rng.uniform(-0.5, 0.5)) if rng.random() < 0.5 else param())
            noisy_ok = False
        elif kind == 6 and others:
            m = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
            qm, _ = np.linalg.qr(m)
            b.matrix(qm, [q, int(rng.choice(others))])
        elif kind == 7:
            m = rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2))
            b.matrix(m / np.linalg.norm(m), [q])
        elif kind == 8:
            terms = []
            for _t in range(3):
                qs = sorted(rng.choice(n, size=int(rng.integers(1, min(n, 3) + 1)), replace=False).tolist())
                terms.append(PauliTerm(complex(rng.uniform(-1, 1)), (), tuple((int(x), "XYZ"[rng.integers(0, 3)]) for x in qs)))
            b.hamevo(PauliSum(n, terms), param())
            noisy_ok = False
        elif kind == 9:
            h = rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2))
            b.hamevo(h + h.conj().T, param(), [q])
            noisy_ok = False
        else:
            b.gate("H", q)
        if noisy_ok and rng.random() < 0.5:
            for _k in range(int(rng.integers(1, 3))):
                name, probs = PROTOCOLS[rng.integers(0, len(PROTOCOLS))]
                b.noise(name, *probs)
    return b.build(), values


@pytest.mark.parametrize("name,probs", PROTOCOLS)
def test_channels_are_trace_preserving(name: str, probs: tuple) -> None:
    ks = density.kraus_operators(name, probs)
    assert np.allclose(sum(k.conj().T @ k for k in ks), np.eye(2), atol=1e-14)
    s = density.channel_superoperator(ks)
    rho = np.array([[0.3, 0.1 - 0.2j], [0.1 + 0.2j, 0.7]])
    assert np.allclose((s @ rho.reshape(4)).reshape(2, 2), sum(k @ rho @ k.conj().T for k in ks), atol=1e-14)


def test_known_answers_of_the_textbook_channels() -> None:
    z = PauliSum(1, [PauliTerm(1.0 + 0j, (), ((0, "Z"),))])
    x = PauliSum(1, [PauliTerm(1.0 + 0j, (), ((0, "X"),))])

    def expect(prog: Program, obs: PauliSum) -> float:
        rho = run_vectorised(prog, {})[0].numpy()
        o = sum(t.const * np.linalg.multi_dot([np.eye(2)] + [CONST_MATS[p] for _, p in t.paulis]) for t in obs.terms)
        return float(np.trace(o @ rho).real)

    p = 0.2
    assert expect(ProgramBuilder(1).X(0).noise("BitFlip", p).build(), z) == pytest.approx(-(1 - 2 * p), abs=1e-14)
    assert expect(ProgramBuilder(1).H(0).noise("PhaseFlip", p).build(), x) == pytest.approx(1 - 2 * p, abs=1e-14)
    assert expect(ProgramBuilder(1).X(0).noise("Depolarizing", p).build(), z) == pytest.approx(-(1 - 4 * p / 3), abs=1e-14)
    assert expect(ProgramBuilder(1).X(0).noise("AmplitudeDamping", p).build(), z) == pytest.approx(-1 + 2 * p, abs=1e-14)
    assert expect(ProgramBuilder(1).H(0).noise("PhaseDamping", p).build(), x) == pytest.approx(np.sqrt(1 - p), abs=1e-14)
    assert expect(ProgramBuilder(1).H(0).noise("PauliChannel", 0.1, 0.2, 0.3).build(), x) == pytest.approx(1 - 2 * (0.2 + 0.3), abs=1e-14)
    # generalised amplitude damping drives towards p|0><0| + (1-p)|1><1|
    g, pp = 1.0, 0.7
    assert expect(ProgramBuilder(1).H(0).noise("GeneralizedAmplitudeDamping", pp, g).build(), z) == pytest.approx(2 * pp - 1, abs=1e-14)


@pytest.mark.parametrize("seed", range(12))
def test_vectorised_program_equals_dense_kraus_evolution(seed: int) -> None:
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 4))
    prog, values = random_noisy_program(rng, n, int(rng.integers(3, 9)))
    B = max([v.shape[0] for v in values.values()] + [1])
    mode = seed % 3
    if mode == 0:
        state, rho0 = None, [np.diag([1.0] + [0.0] * (2**n - 1)).astype(np.complex128)] * B
    elif mode == 1:
        psi = torch.tensor(rng.normal(size=(B, 2**n)) + 1j * rng.normal(size=(B, 2**n)))
        psi = psi / psi.norm(dim=1, keepdim=True)
        state, rho0 = psi, [np.outer(p.numpy(), p.numpy().conj()) for p in psi]
    else:
        a = rng.normal(size=(B, 2**n, 2**n)) + 1j * rng.normal(size=(B, 2**n, 2**n))
        r = a @ a.conj().transpose(0, 2, 1)
        r = r / np.trace(r, axis1=1, axis2=2)[:, None, None]
        state, rho0 = torch.tensor(r).as_subclass(density.DensityMatrix), list(r)
    got = run_vectorised(prog, values, state)
    assert got.shape[0] == B
    for b in range(B):
        want = dense_density(prog, values, rho0[b], b)
        assert np.abs(got[b].numpy() - want).max() < 1e-10


def test_noise_free_program_gives_the_pure_projector() -> None:
    rng = np.random.default_rng(5)
    prog, values = random_noisy_program(rng, 3, 8)
    for op in prog.ops:
        op.noise = ()
    psi = oracle.run(prog, values)
    rho = run_vectorised(prog, values)
    want = torch.einsum("bi,bj->bij", psi, psi.conj())
    assert (rho - want).abs().max() < 1e-10


def test_noisy_programs_are_refused_by_the_state_vector_engine() -> None:
    from qadence_b200 import engine

    prog = ProgramBuilder(2).H(0).noise("BitFlip", 0.1).CNOT(0, 1).build()
    assert prog.is_noisy and Program.loads(prog.dumps()).ops[0].noise == (("BitFlip", (0.1,)),)
    with pytest.raises(ValueError, match="density"):
        engine.CompiledProgram(prog)


def test_readout_confusion_program() -> None:
    rng = np.random.default_rng(0)
    probs = rng.random((2, 8))
    probs /= probs.sum(1, keepdims=True)
    ro = density.ReadoutNoise(3, error_probability=[0.1, 0.0, 0.3])
    out = oracle.run(ro.program(), {}, torch.tensor(probs, dtype=torch.complex128)).real.numpy()
    c = [np.array([[1 - p, p], [p, 1 - p]]) for p in (0.1, 0.0, 0.3)]
    full = np.kron(np.kron(c[0], c[1]), c[2])
    assert np.allclose(out, probs @ full.T, atol=1e-14) and np.allclose(out.sum(1), 1.0)
    cm = rng.random((8, 8))
    cm /= cm.sum(0, keepdims=True)
    out2 = oracle.run(density.ReadoutNoise(3, confusion_matrix=cm).program(), {}, torch.tensor(probs, dtype=torch.complex128)).real.numpy()
    assert np.allclose(out2, probs @ cm.T, atol=1e-14)
