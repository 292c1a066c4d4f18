"""Trailing X / CNOT / SWAP gates absorbed into Pauli-sum observables (`qadence_b200/clifford.py`,
`engine.absorb_trailing_permutation`): P† O P must be the same operator, and the trimmed circuit with the conjugated
observable must give the oracle's expectation values and gradients of the full circuit."""
from __future__ import annotations

import numpy as np
import pytest
import torch

from oracle import statevec as oracle
from qadence_b200 import clifford
from qadence_b200.program import CONST_MATS, Op, OpKind, PauliSum, PauliTerm, Program, ProgramBuilder, hea_program, total_magnetization

PM = {"X": np.array([[0, 1], [1, 0]], dtype=complex), "Y": np.array([[0, -1j], [1j, 0]], dtype=complex), "Z": np.diag([1.0 + 0j, -1])}


def dense_pauli_sum(ps: PauliSum) -> np.ndarray:
    n = ps.n_qubits
    out = np.zeros((2**n, 2**n), dtype=complex)
    for t in ps.terms:
        m = np.ones((1, 1), dtype=complex)
        d = dict(t.paulis)
        for q in range(n):
            m = np.kron(m, PM[d[q]] if q in d else np.eye(2))
        out += complex(t.const) * m
    return out


def dense_unitary(ops, n: int) -> np.ndarray:
    cols = oracle.run(Program(n, list(ops)), {}, torch.eye(2**n, dtype=torch.complex128))
    return cols.numpy().T


def random_permutation_gates(rng, n: int, k: int) -> list[Op]:
    b = ProgramBuilder(n)
    for _ in range(k):
        kind = int(rng.integers(3))
        q = [int(x) for x in rng.permutation(n)]
        if kind == 0:
            b.X(q[0])
        elif kind == 1:
            b.CNOT(q[0], q[1])
        else:
            b.SWAP(q[0], q[1])
    return b.build().ops


def random_pauli_sum(rng, n: int, terms: int) -> PauliSum:
    out = []
    for _ in range(terms):
        qs = sorted(int(x) for x in rng.choice(n, size=int(rng.integers(1, n + 1)), replace=False))
        out.append(PauliTerm(float(rng.normal()), (), tuple((q, "XYZ"[int(rng.integers(3))]) for q in qs)))
    return PauliSum(n, out)


@pytest.mark.parametrize("seed", range(8))
def test_conjugation_is_the_same_operator(seed: int) -> None:
    rng = np.random.default_rng(seed)
    n = 4
    gates = random_permutation_gates(rng, n, 9)
    obs = random_pauli_sum(rng, n, 6)
    P = dense_unitary(gates, n)
    want = P.conj().T @ dense_pauli_sum(obs) @ P
    got = dense_pauli_sum(clifford.conjugate_observable(obs, gates))
    assert np.abs(got - want).max() < 1e-13
    assert len(clifford.conjugate_observable(obs, gates).terms) == len(obs.terms)


def test_split_takes_only_closing_permutation_gates() -> None:
    b = ProgramBuilder(3)
    b.RX(0, "a").CNOT(0, 1).RZ(2, "b").CNOT(1, 2).SWAP(0, 2).X(1)
    prog = b.build()
    body, gates = clifford.split_trailing_permutation(prog)
    assert len(gates) == 3 and len(body.ops) == 3 and body.ops[-1].kind == OpKind.RZ
    b2 = ProgramBuilder(3)
    b2.RX(0, "a").CRZ(0, 1, "c")
    assert clifford.split_trailing_permutation(b2.build())[1] == []
    toffoli = Op(OpKind.MAT1, (2,), (0, 1), matrix=CONST_MATS["X"].copy(), label="Toffoli")
    assert not clifford.is_permutation_gate(toffoli)  # a permutation, but not a Clifford gate: Pauli strings do not stay strings


def test_hea_ladder_turns_total_magnetization_into_prefix_parities() -> None:
    n = 6
    prog = hea_program(n, 2)
    body, gates = clifford.split_trailing_permutation(prog)
    assert len(gates) == n - 1 and all(g.controls for g in gates)
    conj = clifford.conjugate_observable(total_magnetization(n), gates)
    assert all(all(p == "Z" for _, p in t.paulis) for t in conj.terms)
    rng = np.random.default_rng(0)
    values = {nm: torch.tensor(rng.uniform(0, 2 * np.pi, size=3)) for nm in prog.param_names}
    full = oracle.expectation(prog, [total_magnetization(n)], values)
    trimmed = oracle.expectation(body, [conj], values)
    assert (full - trimmed).abs().max() < 1e-13


def test_engine_absorbs_only_when_it_pays(monkeypatch) -> None:
    from qadence_b200 import cabi, engine

    try:
        cabi.load()
    except cabi.B200svError:
        pytest.skip("libb200sv.so not built")
    # hea(20, 3): the closing ladder costs a sweep of its own (4 -> 3) and Z_tot becomes 20 CONSTANT prefix-parity strings, which the
    # Pauli kernels read from a table: absorbed
    big = engine.CompiledProgram(hea_program(20, 3))
    cob = engine.CompiledObservable(total_magnetization(20))
    body, cobs = engine.absorb_trailing_permutation(big, [cob])
    assert body is not big and len(body.program.ops) == len(big.program.ops) - 19 and body.names == big.names
    assert cobs[0] is not cob and len(cobs[0].obs.terms) == 20 and cobs[0].pauli.table_split() is not None
    assert engine.absorb_trailing_permutation(big, [cob])[1][0] is cobs[0]  # cached
    # cfg2 (depth 8): the planner folds the ladder into the other sweeps (8 sweeps either way): nothing to gain, Z_tot stays single-Z
    cfg2 = engine.CompiledProgram(hea_program(20, 8))
    same, c2 = engine.absorb_trailing_permutation(cfg2, [cob])
    assert same is cfg2 and c2[0] is cob
    # the same ladder under an observable with PARAMETRIC coefficients would put 19 strings on the slow path: not absorbed ...
    pobs = engine.CompiledObservable(PauliSum(20, [PauliTerm(1.0, (f"w{q}",), ((q, "Z"),)) for q in range(20)]))
    same, pc = engine.absorb_trailing_permutation(big, [pobs])
    assert same is big and pc[0] is pobs
    # ... unless forced
    monkeypatch.setenv("B200SV_ABSORB_PERM", "force")
    big2 = engine.CompiledProgram(hea_program(20, 3))
    body2, pc2 = engine.absorb_trailing_permutation(big2, [pobs])
    assert body2 is not big2 and pc2[0] is not pobs
    monkeypatch.delenv("B200SV_ABSORB_PERM")
    # a closing layer of long-range SWAPs costs sweeps and leaves Z_tot a sum of single Z: absorbed
    b = ProgramBuilder(20)
    for op in hea_program(20, 2).ops:
        b._add(op)
    for q in range(20):
        b.RX(q, f"last{q}")
    for q in range(10):
        b.SWAP(q, 19 - q)
    sw = engine.CompiledProgram(b.build())
    body, cobs = engine.absorb_trailing_permutation(sw, [cob])
    assert body is not sw and all(len(t.paulis) == 1 for t in cobs[0].obs.terms)
    small = engine.CompiledProgram(hea_program(6, 2))  # one sweep either way: nothing to gain
    c6 = engine.CompiledObservable(total_magnetization(6))
    same, cobs6 = engine.absorb_trailing_permutation(small, [c6])
    assert same is small and cobs6[0] is c6
    # gradients of the trimmed problem equal the full one (oracle, a smaller instance of the same structure)
    prog = hea_program(5, 2)
    bodyp, gates = clifford.split_trailing_permutation(prog)
    obs = PauliSum(5, [PauliTerm(1.0, (), ((q, "Z"),)) for q in range(5)] + [PauliTerm(0.4, (), ((0, "X"), (3, "Y")))])
    rng = np.random.default_rng(1)
    values = {nm: torch.tensor(rng.uniform(0, 2 * np.pi, size=2)) for nm in prog.param_names}
    e1, g1 = oracle.adjoint_expectation_and_grads(prog, obs, values)
    e2, g2 = oracle.adjoint_expectation_and_grads(bodyp, clifford.conjugate_observable(obs, gates), values)
    assert (e1 - e2).abs().max() < 1e-13 and max((g1[k] - g2[k]).abs().max() for k in g1) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.complex128, torch.complex64])
def test_gpu_expectation_with_and_without_absorption(monkeypatch, dtype) -> None:
    """hea(18, 3) on the GPU: the closing ladder is absorbed (it saves a sweep at 2^10-amplitude tiles: 4 -> 3); values and all
    gradients equal the oracle's for the FULL circuit, and the run with `B200SV_ABSORB_PERM=0`."""
    from qadence_b200 import engine

    n, B = 18, 1
    prog = hea_program(n, 3)
    obs = PauliSum(n, [PauliTerm(1.0, (), ((q, "Z"),)) for q in range(n)] + [PauliTerm(0.5, (), ((1, "X"), (9, "Y"))), PauliTerm(-0.3, (), ((12, "Y"),))])
    rng = np.random.default_rng(4)
    base = {nm: torch.tensor(rng.uniform(0, 2 * np.pi, size=B)) for nm in prog.param_names}
    want_e, want_g = oracle.adjoint_expectation_and_grads(prog, obs, base)
    tol = 1e-10 if dtype == torch.complex128 else 2e-4
    results = {}
    for flag in ("force", "0"):
        monkeypatch.setenv("B200SV_ABSORB_PERM", flag)
        cp = engine.CompiledProgram(prog)
        cob = engine.CompiledObservable(obs)
        body, _ = engine.absorb_trailing_permutation(cp, [cob])
        assert (body is not cp) == (flag == "force")
        values = {k: v.clone().requires_grad_(True) for k, v in base.items()}
        e = engine.expectation(cp, [cob], values, dtype=dtype, device="cuda:0")
        e.sum().backward()
        assert (e[:, 0].detach().cpu().double() - want_e).abs().max() < tol
        assert max((values[k].grad.cpu() - want_g[k]).abs().max() for k in values) < tol
        results[flag] = e.detach().cpu()
    assert (results["force"] - results["0"]).abs().max() < tol


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.complex128, torch.complex64])
def test_gpu_tabulated_diagonal_observables(monkeypatch, dtype) -> None:
    """Observables with constant ZZ.. strings on a state of >= 2^22 amplitudes take the tabulated-diagonal kernels
    (`b200sv_pauli_expectation_diag`, `b200sv_pauli_apply_diag` with a per-batch scale): expectation values and gradients of TWO
    observables with different cotangents equal the oracle's and the term-by-term kernels' (`B200SV_DIAG_TABLE=0`)."""
    from qadence_b200 import engine

    n, B = 11, 2048  # 2^22 amplitudes
    prog = hea_program(n, 1)
    ising = PauliSum(n, [PauliTerm(0.7 + 0.1 * q, (), ((q, "Z"), (q + 1, "Z"))) for q in range(n - 1)] + [PauliTerm(0.3, (), ((q, "Z"),)) for q in range(n)]
                     + [PauliTerm(0.25, (), ((0, "X"), (5, "X"))), PauliTerm(-0.4, (), ((2, "Y"),))])
    parity = PauliSum(n, [PauliTerm(1.0, (), tuple((q, "Z") for q in range(k + 1))) for k in range(n)])
    rng = np.random.default_rng(8)
    base = {nm: torch.tensor(rng.uniform(0, 2 * np.pi, size=B)) for nm in prog.param_names}
    wts = torch.tensor(rng.normal(size=(B, 2)))
    tol = 1e-10 if dtype == torch.complex128 else 3e-4
    cols = [0, 1, B - 1]
    sub = {k: v[cols] for k, v in base.items()}
    want = oracle.expectation(prog, [ising, parity], sub)
    out = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("B200SV_DIAG_TABLE", flag)
        cp = engine.CompiledProgram(prog)
        cobs = [engine.CompiledObservable(ising), engine.CompiledObservable(parity)]
        values = {k: v.clone().requires_grad_(True) for k, v in base.items()}
        e = engine.expectation(cp, cobs, values, dtype=dtype, device="cuda:0")
        psi_probe = torch.empty(B, 1 << n, dtype=dtype, device="cuda:0")
        assert cobs[0].uses_table(psi_probe) == (flag == "1")
        (e.cpu().double() * wts).sum().backward()
        assert (e[cols].detach().cpu().double() - want).abs().max() < tol
        out[flag] = (e.detach().cpu().double(), torch.stack([values[k].grad for k in values]))
    assert (out["1"][0] - out["0"][0]).abs().max() < tol
    assert (out["1"][1] - out["0"][1]).abs().max() < 10 * tol
    # gradients against the oracle for the sampled batch elements
    g_want = {k: 0 for k in base}
    for o, obs in enumerate((ising, parity)):
        _, g = oracle.adjoint_expectation_and_grads(prog, obs, sub)
        for k in g:
            g_want[k] = g_want[k] + g[k] * wts[cols, o]
    got = {k: out["1"][1][i][cols] for i, k in enumerate(base)}
    assert max((got[k] - g_want[k]).abs().max() for k in base) < 10 * tol
