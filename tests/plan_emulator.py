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


def _perm_src(g, tile, idx: np.ndarray, gidx: np.ndarray | None = None) -> np.ndarray:
    """Index map of one X / CNOT / SWAP record (self-inverse): new[i] = old[src[i]]."""
    gidx = idx if gidx is None else gidx
    ta = tile[int(g["a"])]
    if int(g["kind"]) == P.GK_SWAP:
        tb = tile[int(g["a2"])]
        bit_a = (idx >> np.uint64(ta)) & np.uint64(1)
        bit_b = (idx >> np.uint64(tb)) & np.uint64(1)
        return (idx ^ ((bit_a ^ bit_b) * np.uint64((1 << ta) | (1 << tb)))).astype(np.int64)
    c = int(g["ext_bit"])
    if c < 0:
        return (idx ^ np.uint64(1 << ta)).astype(np.int64)
    cb = (gidx >> np.uint64(c)) & np.uint64(1)
    return (idx ^ (cb * np.uint64(1 << ta))).astype(np.int64)


def _apply_perm_record(st: np.ndarray, g, tile, idx: np.ndarray, gidx: np.ndarray | None = None) -> np.ndarray:
    return st[:, _perm_src(g, tile, idx, gidx)]


def _check_perm_map(plan: P.SweepPlan, sw, rd, tile, n: int) -> None:
    """The round's packed affine map must equal the composition of its permutation records."""
    m = len(tile)
    N = 1 << n
    idx = np.arange(N, dtype=np.uint64)
    recs = plan.gates[int(rd["first_gate"]) : int(rd["first_gate"]) + int(rd["n_gates"])]
    pos = idx.copy()  # where the amplitude initially at i ends up
    for g in recs:
        if int(g["xkind"]) == P.XK_PERM:
            pos = _perm_src(g, tile, pos).astype(np.uint64)  # each gate is an involution: dest = src map
    # packed map on tile-local indices
    local = np.zeros(N, dtype=np.uint64)
    for li, gb in enumerate(tile):
        local |= ((idx >> np.uint64(gb)) & np.uint64(1)) << np.uint64(li)
    img = np.full(N, int(rd["pconst"]), dtype=np.uint64)
    for li in range(m):
        img ^= ((local >> np.uint64(li)) & np.uint64(1)) * np.uint64(int(rd["pcol"][li]))
    for k in range(int(rd["n_pext"])):
        bit, vec = int(rd["pext"][k]["bit"]), int(rd["pext"][k]["vec"])
        assert bit not in tile
        img ^= ((idx >> np.uint64(bit)) & np.uint64(1)) * np.uint64(vec)
    want_local = np.zeros(N, dtype=np.uint64)
    for li, gb in enumerate(tile):
        want_local |= ((pos >> np.uint64(gb)) & np.uint64(1)) << np.uint64(li)
    outer_mask = np.uint64(sum(1 << b for b in range(n) if b not in tile))
    assert np.array_equal(pos & outer_mask, idx & outer_mask), "permutation must stay inside the tile"
    assert np.array_equal(img, want_local), "packed affine map != composed permutation gates"
    is_id = np.array_equal(img, local)
    if int(rd["n_pext"]) == 0:
        assert bool(int(rd["flags"]) & P.RF_PERM) == (not is_id)
    else:  # a map that depends on outer / rank bits may reduce to the identity on the indices seen here
        assert int(rd["flags"]) & P.RF_PERM
    # M-slot bookkeeping
    r = int(sw["r"])
    for a in range(4):
        ms = int(rd["mslot"][a])
        if ms >= 0:
            assert a < r
            leaders = [g for g in recs if int(g["chain"]) >= 1 and int(g["xidx"]) == ms]
            assert len(leaders) == 1 and int(leaders[0]["a"]) == a and (int(leaders[0]["flags"]) & P.GF_MSLOT)
            assert int(leaders[0]["ctrl_ext"]) == 0 and int(leaders[0]["ctrl_reg"]) == 0


def emulate(plan: P.SweepPlan, state: np.ndarray, params: np.ndarray, dagger: bool = False, index_hi: int = 0) -> np.ndarray:
    """state [B, 2^n] complex → new state after all sweeps (forward, or reversed with daggers)."""
    st = state.astype(np.complex128).copy()
    B, N = st.shape
    n = plan.n_qubits
    idx = np.arange(N, dtype=np.uint64)
    gidx = idx | np.uint64(index_hi)  # index seen by control / diagonal predicates (rank bits of a shard)
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
            _check_perm_map(plan, sw, rd, tile, n)
            for gi in (reversed(gl) if dagger else gl):
                g = plan.gates[gi]
                if int(g["xkind"]) == P.XK_PERM:
                    st = _apply_perm_record(st, g, tile, idx, gidx)
                    continue
                cext = int(g["ctrl_ext"])
                assert cext & regmask == 0
                cmask = cext
                for a in range(r):
                    if (int(g["ctrl_reg"]) >> a) & 1:
                        cmask |= 1 << regbit[a]
                active = (gidx & np.uint64(cmask)) == np.uint64(cmask)
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
                    bit = ((gidx >> np.uint64(tbit)) & np.uint64(1)).astype(np.int64)
                    if tbit >= n:  # external diagonal on a rank bit
                        new = mat[bit, bit] * v
                    else:
                        partner = (idx ^ np.uint64(1 << tbit)).astype(np.int64)
                        new = mat[bit, bit] * v + mat[bit, 1 - bit] * v[partner]
                    st[b] = np.where(active, new, v)
    return st


# ---------------------------------------------------------------------------------------------------
# chain-merged execution + adjoint moments, as sweep.cuh's resolve_gate / chain_gradients do it
# ---------------------------------------------------------------------------------------------------
_SIG = {
    "I": np.eye(2, dtype=np.complex128),
    "X": np.array([[0, 1], [1, 0]], dtype=np.complex128),
    "Y": np.array([[0, -1j], [1j, 0]], dtype=np.complex128),
    "Z": np.array([[1, 0], [0, -1]], dtype=np.complex128),
}


def _apply_1q(v: np.ndarray, mat: np.ndarray, tbit: int, active: np.ndarray, idx: np.ndarray) -> np.ndarray:
    bit = ((idx >> np.uint64(tbit)) & np.uint64(1)).astype(np.int64)
    partner = (idx ^ np.uint64(1 << tbit)).astype(np.int64)
    new = mat[bit, bit] * v + mat[bit, 1 - bit] * v[partner]
    return np.where(active, new, v)


def emulate_adjoint(plan: P.SweepPlan, psi: np.ndarray, lam: np.ndarray, params: np.ndarray):
    """Backward walk over exec entries (chains merged): returns (psi0, lam0, grads[n_slots, B])."""
    psi = psi.astype(np.complex128).copy()
    lam = lam.astype(np.complex128).copy()
    B, N = psi.shape
    grads = np.zeros_like(params, dtype=np.float64)
    idx = np.arange(N, dtype=np.uint64)
    for s in reversed(range(plan.n_sweeps)):
        sw = plan.sweeps[s]
        m, r = int(sw["m"]), int(sw["r"])
        tile = [int(x) for x in sw["tile_bits"][:m]]
        for ri in reversed(range(int(sw["first_round"]), int(sw["first_round"]) + int(sw["n_rounds"]))):
            rd = plan.rounds[ri]
            regbit = [tile[int(p)] for p in rd["regpos"][:r]]
            recs = plan.gates[int(rd["first_gate"]) : int(rd["first_gate"]) + int(rd["n_gates"])]
            for g in reversed(recs):  # P phase first when walking backwards
                if int(g["xkind"]) == P.XK_PERM:
                    src = _perm_src(g, tile, idx)
                    psi, lam = psi[:, src], lam[:, src]
            leaders = [i for i, g in enumerate(recs) if g["chain"] >= 1]
            assert len(leaders) == int(rd["n_exec"]) + sum(1 for a in range(4) if int(rd["mslot"][a]) >= 0)
            for li in reversed(leaders):
                g = recs[li]
                k = int(g["chain"])
                chain = recs[li : li + k]
                assert all(int(c["chain"]) == 0 for c in chain[1:]) and all(int(c["xidx"]) == int(g["xidx"]) for c in chain)
                cmask = int(g["ctrl_ext"])
                for a in range(r):
                    if (int(g["ctrl_reg"]) >> a) & 1:
                        cmask |= 1 << regbit[a]
                active = (idx & np.uint64(cmask)) == np.uint64(cmask)
                xk = int(g["xkind"])
                for b in range(B):
                    if xk == P.XK_SCALE:
                        sval = _gate_matrix(plan, g, params, b, False)[1]
                        psi[b] = psi[b] / sval
                        if g["flags"] & P.GF_GRAD:
                            grads[g["grad_row"], b] += 2 * np.real(np.vdot(lam[b], psi[b]))
                        lam[b] = lam[b] * np.conj(sval)
                        continue
                    if xk == P.XK_SWAP:
                        ba, bb = regbit[int(g["a"])], regbit[int(g["a2"])]
                        bit_a = (idx >> np.uint64(ba)) & np.uint64(1)
                        bit_b = (idx >> np.uint64(bb)) & np.uint64(1)
                        src = (idx ^ ((bit_a ^ bit_b) * np.uint64((1 << ba) | (1 << bb)))).astype(np.int64)
                        psi[b] = np.where(active, psi[b][src], psi[b])
                        lam[b] = np.where(active, lam[b][src], lam[b])
                        continue
                    a = int(g["a"])
                    tbit = regbit[a] if a >= 0 else int(g["ext_bit"])
                    mats = [_gate_matrix(plan, c, params, b, False)[1] for c in chain]
                    prod = np.eye(2, dtype=np.complex128)
                    for mt in mats:
                        prod = mt @ prod
                    if xk == P.XK_DIAG:
                        assert prod[0, 1] == 0 and prod[1, 0] == 0
                    if xk == P.XK_REAL:
                        assert np.all(prod.imag == 0)
                    if xk == P.XK_RX:
                        assert abs(prod[0, 0] - prod[1, 1]) < 1e-15 and abs(prod[0, 1] - prod[1, 0]) < 1e-15 and abs(prod[0, 1].real) < 1e-15
                    if any(c["flags"] & P.GF_GRAD for c in chain):
                        # four moments on the POST-chain states, control-active subspace
                        mom = {}
                        for nm, sg in _SIG.items():
                            sp = _apply_1q(psi[b], sg, tbit, active, idx)
                            mom[nm] = np.imag(np.vdot(lam[b][active], sp[active]))
                        W = np.eye(2, dtype=np.complex128)
                        for i in range(k - 1, -1, -1):
                            c = chain[i]
                            if c["flags"] & P.GF_GRAD:
                                kind = int(c["kind"])
                                Pm = {P.GK_RX: _SIG["X"], P.GK_RY: _SIG["Y"], P.GK_RZ: _SIG["Z"], P.GK_PHASE: _SIG["Z"]}[kind]
                                cI = -1.0 if kind == P.GK_PHASE else 0.0
                                A = W @ Pm @ W.conj().T
                                az, ax, ay = A[0, 0].real, A[1, 0].real, A[1, 0].imag
                                grads[c["grad_row"], b] += cI * mom["I"] + ax * mom["X"] + ay * mom["Y"] + az * mom["Z"]
                            W = W @ mats[i]
                    dag = prod.conj().T
                    psi[b] = _apply_1q(psi[b], dag, tbit, active, idx)
                    lam[b] = _apply_1q(lam[b], dag, tbit, active, idx)
    return psi, lam, grads
