"""Trailing permutation gates of a circuit, absorbed into Pauli-sum observables.

A circuit U = P·V whose last gates P are basis-state permutations — X, CNOT, SWAP: the closing entangler layer of every
hardware-efficient ansatz (`/root/reference/qadence/constructors/hea.py`) — has ⟨0|U† O U|0⟩ = ⟨0|V† (P† O P) V|0⟩, and P† O P is
again a Pauli sum with the same number of terms (Clifford conjugation: every string maps to ± one string).  The expectation
path therefore runs V only and measures the conjugated observable: one full state round trip less per direction (for cfg2 the
ladder of the last layer does not fit the tile of the last sweep and used to cost a forward sweep of 2·S and an adjoint sweep of
4·S bytes on its own).  Gate parameters all live in V, so gradients are unchanged; `run` and `sample` still apply P.

The single- and two-qubit conjugation tables are computed from the 2×2 / 4×4 matrices at import, not written down by hand.
"""
from __future__ import annotations

import itertools
from typing import Sequence

import numpy as np

from .program import Op, OpKind, PauliSum, PauliTerm, Program

_P = {"I": np.eye(2, dtype=np.complex128), "X": np.array([[0, 1], [1, 0]], dtype=np.complex128),
      "Y": np.array([[0, -1j], [1j, 0]], dtype=np.complex128), "Z": np.array([[1, 0], [0, -1]], dtype=np.complex128)}
_CNOT = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=np.complex128)  # control = first (MSB) factor


def _cnot_table() -> dict[tuple[str, str], tuple[str, str, int]]:
    """(P_control, P_target) -> (P'_control, P'_target, sign) with CNOT (P_c ⊗ P_t) CNOT = sign · P'_c ⊗ P'_t."""
    out = {}
    for a, b in itertools.product("IXYZ", repeat=2):
        m = _CNOT @ np.kron(_P[a], _P[b]) @ _CNOT
        for a2, b2 in itertools.product("IXYZ", repeat=2):
            ref = np.kron(_P[a2], _P[b2])
            for sign in (1, -1):
                if np.array_equal(m, sign * ref):
                    out[(a, b)] = (a2, b2, sign)
        assert (a, b) in out
    return out


CNOT_TABLE = _cnot_table()
X_SIGN = {"I": 1, "X": 1, "Y": -1, "Z": -1}  # X P X = sign · P


def is_permutation_gate(op: Op) -> bool:
    if op.noise:
        return False
    if op.kind == OpKind.SWAP:
        return not op.controls
    if op.kind == OpKind.MAT1 and len(op.controls) <= 1 and op.matrix is not None:
        return bool(np.array_equal(np.asarray(op.matrix, dtype=np.complex128).reshape(2, 2), _P["X"]))
    return False


def split_trailing_permutation(program: Program) -> tuple[Program, list[Op]]:
    """(V, [g_1 … g_k]) with program = g_k ⋯ g_1 · V and every g an X / CNOT / SWAP; k = 0 if the circuit does not end with one."""
    ops = list(program.ops)
    k = len(ops)
    while k > 0 and is_permutation_gate(ops[k - 1]):
        k -= 1
    return Program(program.n_qubits, ops[:k]), ops[k:]


def conjugate_term(term: PauliTerm, gates: Sequence[Op]) -> PauliTerm:
    """P† · term · P for P = g_k ⋯ g_1 (`gates` in circuit order): conjugate by g_k first, then g_{k−1}, …"""
    paulis = dict(term.paulis)
    sign = 1
    for g in reversed(gates):
        if g.kind == OpKind.SWAP:
            a, b = g.targets
            pa, pb = paulis.pop(a, None), paulis.pop(b, None)
            if pa is not None:
                paulis[b] = pa
            if pb is not None:
                paulis[a] = pb
        elif g.controls:
            c, t = g.controls[0], g.targets[0]
            pc, pt, s_ = CNOT_TABLE[(paulis.get(c, "I"), paulis.get(t, "I"))]
            sign *= s_
            for q, p in ((c, pc), (t, pt)):
                if p == "I":
                    paulis.pop(q, None)
                else:
                    paulis[q] = p
        else:
            sign *= X_SIGN[paulis.get(g.targets[0], "I")]
    return PauliTerm(term.const * sign, term.names, tuple(sorted(paulis.items())), term.expr)


def conjugate_observable(obs: PauliSum, gates: Sequence[Op]) -> PauliSum:
    return PauliSum(obs.n_qubits, [conjugate_term(t, gates) for t in obs.terms])
