"""CPU emulation of what `sweep_kernel` does with a packed `SweepPlan` (test infrastructure).

It interprets the binary records (sweeps → rounds → gates; register indices, control masks,
external diagonal bits, constant pool) exactly as `qadence_b200/csrc/sweep.cuh` documents them, but
on a numpy state, so the planner (`qadence_b200/plan.py`) can be validated against the oracle
without a GPU.  The real kernel is validated against the oracle on the GPU (tests/test_gpu_*.py).
"""

from __future__ import annotations

import numpy as np

from qadence_b200 import plan as P


def _consts(plan: P.SweepPlan, cidx: int, k: int) -> np.ndarray:
    c = plan.consts[cidx : cidx + k]
    return c[:, 0] + 1j * c[:, 1]


def _gate_matrix(plan: P.SweepPlan, g, params: np.ndarray, b: int, dagger: bool):
    """Return ('mat', 2x2) | ('swap',) | ('scale', s)."""
    kind = int(g["kind"])
    if kind in (P.GK_RX, P.GK_RY, P.GK_RZ, P.GK_PHASE, P.GK_SCALE):
        th = params[g["slot"], b] if g["slot"] >= 0 else _consts(plan, g["cidx"], 1)[0]
    if kind == P.GK_MAT:
        m = _consts(plan, g["cidx"], 4).reshape(2, 2)
    elif kind == P.GK_REAL:
        c = plan.consts[g["cidx"] : g["cidx"] + 2].reshape(-1)
        m = np.array([[c[0], c[1]], [c[2], c[3]]], dtype=np.complex128)
    elif kind == P.GK_RX:
        th = np.real(th)
        m = np.array([[np.cos(th / 2), -1j * np.sin(th / 2)], [-1j * np.sin(th / 2), np.cos(th / 2)]])
    elif kind == P.GK_RY:
        th = np.real(th)
        m = np.array([[np.cos(th / 2), -np.sin(th / 2)], [np.sin(th / 2), np.cos(th / 2)]], dtype=np.complex128)
    elif kind == P.GK_RZ:
        th = np.real(th)
        m = np.diag([np.exp(-0.5j * th), np.exp(0.5j * th)])
    elif kind == P.GK_PHASE:
        th = np.real(th)
        m = np.diag([1.0, np.exp(1j * th)])
    elif kind == P.GK_DIAG:
        d = _consts(plan, g["cidx"], 2)
        m = np.diag(d)
    elif kind == P.GK_X:
        m = np.array([[0, 1], [1, 0]], dtype=np.complex128)
    elif kind == P.GK_SWAP:
        return ("swap",)
    elif kind == P.GK_SCALE:
        return ("scale", np.conj(th) if dagger else th)
    else:
        raise ValueError(kind)
    return ("mat", m.conj().T if dagger else m)


def emulate(plan: P.SweepPlan, state: np.ndarray, params: np.ndarray, dagger: bool = False) -> np.ndarray:
    """state [B, 2^n] complex → new state after all sweeps (forward, or reversed with daggers)."""
    st = state.astype(np.complex128).copy()
    B, N = st.shape
    n = plan.n_qubits
    idx = np.arange(N, dtype=np.uint64)
    sweeps = list(range(plan.n_sweeps))
    for s in (reversed(sweeps) if dagger else sweeps):
        sw = plan.sweeps[s]
        m, r = int(sw["m"]), int(sw["r"])
        tile = [int(x) for x in sw["tile_bits"][:m]]
        assert len(set(tile)) == m and all(0 <= t < n for t in tile)
        outer = [int(x) for x in sw["outer_bits"][: int(sw["n_outer"])]]
        assert sorted(tile + outer) == list(range(n))
        rnds = list(range(int(sw["first_round"]), int(sw["first_round"]) + int(sw["n_rounds"])))
        for ri in (reversed(rnds) if dagger else rnds):
            rd = plan.rounds[ri]
            regpos = [int(x) for x in rd["regpos"][:r]]
            thrpos = [int(x) for x in rd["thrpos"][: m - r]]
            assert sorted(regpos + thrpos) == list(range(m)), (regpos, thrpos, m)
            regbit = [tile[p] for p in regpos]  # global bit position of register index a
            regmask = sum(1 << b for b in regbit)
            gl = list(range(int(rd["first_gate"]), int(rd["first_gate"]) + int(rd["n_gates"])))
            for gi in (reversed(gl) if dagger else gl):
                g = plan.gates[gi]
                cext = int(g["ctrl_ext"])
                assert cext & regmask == 0
                cmask = cext
                for a in range(r):
                    if (int(g["ctrl_reg"]) >> a) & 1:
                        cmask |= 1 << regbit[a]
                active = (idx & np.uint64(cmask)) == np.uint64(cmask)
                for b in range(B):
                    what = _gate_matrix(plan, g, params, b, dagger)
                    v = st[b]
                    if what[0] == "scale":
                        st[b] = v * what[1]
                        continue
                    if what[0] == "swap":
                        ba, bb = regbit[int(g["a"])], regbit[int(g["a2"])]
                        bit_a = (idx >> np.uint64(ba)) & np.uint64(1)
                        bit_b = (idx >> np.uint64(bb)) & np.uint64(1)
                        src = idx ^ ((bit_a ^ bit_b) * np.uint64((1 << ba) | (1 << bb)))
                        new = np.where(active, v[src.astype(np.int64)], v)
                        st[b] = new
                        continue
                    mat = what[1]
                    a = int(g["a"])
                    tbit = regbit[a] if a >= 0 else int(g["ext_bit"])
                    if a < 0:
                        assert mat[0, 1] == 0 and mat[1, 0] == 0, "external targets must be diagonal"
                    bit = ((idx >> np.uint64(tbit)) & np.uint64(1)).astype(np.int64)
                    partner = (idx ^ np.uint64(1 << tbit)).astype(np.int64)
                    new = mat[bit, bit] * v + mat[bit, 1 - bit] * v[partner]
                    st[b] = np.where(active, new, v)
    return st
