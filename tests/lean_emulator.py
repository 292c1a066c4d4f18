"""CPU emulation of `lean_sweep_kernel` (qadence_b200/csrc/lean.cuh) driven by the SAME host-built tables the kernel reads
(test infrastructure).

It follows the kernel step by step on a numpy "shared memory": tile staging in the TMA (plain) or gathered (swizzled)
layout, the per-round load / store byte offsets from the lean tables (thread nibbles, register bits, external-control
constants, pivot flips), the normalised 2x2 with its pending scalar, the scaled / sign-flipped adjoint moments — so
the planner's tables and the normalisation algebra are validated against the oracle without a GPU.  The kernel itself
is validated against the oracle on the GPU (tests/test_gpu_*.py).
"""

from __future__ import annotations

import numpy as np

from qadence_b200 import plan as P

from tests.plan_emulator import _SIG, _gate_matrix


def _chain_product(plan: P.SweepPlan, recs, params: np.ndarray, b: int):
    mats = [_gate_matrix(plan, c, params, b, False)[1] for c in recs]
    prod = np.eye(2, dtype=np.complex128)
    for mt in mats:
        prod = mt @ prod
    return prod, mats


def resolve_lean(plan: P.SweepPlan, si: int, params: np.ndarray, b: int, forward: bool):
    """What `resolve_lean_kernel` computes for (sweep, batch element): per entry (beta, gamma, delta), mscale, flips, and
    the sweep scalar."""
    sw = plan.sweeps[si]
    r, fx = int(sw["r"]), int(sw["first_exec"])
    acc = 1.0 + 0.0j
    out = {}
    rounds = range(int(sw["n_rounds"])) if forward else reversed(range(int(sw["n_rounds"])))
    for rr in rounds:
        rd = plan.rounds[int(sw["first_round"]) + rr]
        ms = [int(x) for x in rd["mslot"]] + [int(rd["mslot4"])]
        round_scale = abs(acc) ** 2  # the kernel takes all moments of a round before un-applying any of its slots
        for a in (range(r) if forward else reversed(range(r))):
            x = ms[a]
            if x < 0:
                continue
            li = int(plan.exec_leader[fx + x])
            k = int(plan.gates[li]["chain"])
            recs = plan.gates[li : li + k]
            prod, _ = _chain_product(plan, recs, params, b)
            if not forward:
                prod = prod.conj().T
            mags = np.abs(prod.reshape(-1)) ** 2
            best = 0
            for e in range(4):  # first strict maximum, as in the kernel
                if mags[e] > mags[best]:
                    best = e
            rf, cf = best >> 1, best & 1
            nu = prod[rf, cf]
            E = np.array([[prod[i ^ rf, j ^ cf] for j in range(2)] for i in range(2)])
            N = E / nu if nu != 0 else np.zeros((2, 2), dtype=np.complex128)
            out[x] = dict(beta=N[0, 1], gamma=N[1, 0], delta=N[1, 1], mscale=round_scale, cf=cf, rf=rf,
                          grad=(not forward) and any(int(c["flags"]) & P.GF_GRAD for c in recs), recs=recs)
            acc = acc * nu
    return out, acc


def emulate_lean_sweep(plan: P.SweepPlan, si: int, psi: np.ndarray, lam: np.ndarray | None, params: np.ndarray,
                       forward: bool, grads: np.ndarray | None = None, index_hi: int = 0):
    """One sweep, in place on psi (and lam).  Only for sweeps the planner marked lean."""
    sw = plan.sweeps[si]
    mode = int(sw["io_mode"])
    assert mode != P.IO_GENERIC
    m, r, n = int(sw["m"]), int(sw["r"]), plan.n_qubits
    c128 = plan.cfg.c128
    ebytes = 16 if c128 else 8
    NT, NV, TILE = 1 << (m - r), 1 << r, 1 << m
    tile_bits = [int(x) for x in sw["tile_bits"][:m]]
    outer = [int(x) for x in sw["outer_bits"][: int(sw["n_outer"])]]
    tab = plan.lean_tab_fwd if forward else plan.lean_tab_bwd
    fr, nr = int(sw["first_round"]), int(sw["n_rounds"])
    B = psi.shape[0]
    e_idx = np.arange(TILE)
    dep = np.zeros(TILE, dtype=np.int64)  # tile-local index -> offset in the state
    for i, gb in enumerate(tile_bits):
        dep |= ((e_idx >> i) & 1) << gb
    plain_io = mode == P.IO_LEAN_TMA
    lay = e_idx if plain_io else np.array([P._swz_b(int(e), c128) for e in e_idx])
    if plain_io:  # the TMA box must enumerate the tile in ascending tile-bit order and walk the outer bits with the counter
        tm = plan.tma[si]
        assert int(tm["rank"]) & 0xFF >= 2
        if int(tm["rank"]) & P.TMA_SWIZZLE_FLAG:  # hardware swizzle of the TMA-written tile
            lay = np.array([P._swz_a(int(e), c128) for e in e_idx])
    tid = np.arange(NT)
    jj = np.arange(NV)
    for b in range(B):
        gates, sscal = resolve_lean(plan, si, params, b, forward)
        for c in range(1 << len(outer)):
            base = 0
            for k, ob in enumerate(outer):
                base |= ((c >> k) & 1) << ob
            cta_index = base | index_hi
            if plain_io:
                _check_tma_coords(plan.tma[si], c, b, base, n, tile_bits)
            smem_p = np.zeros(TILE, dtype=np.complex128)
            smem_l = np.zeros(TILE, dtype=np.complex128)
            smem_p[lay] = psi[b, base + dep]
            if lam is not None:
                smem_l[lay] = lam[b, base + dep]
            for rr in range(nr):
                rloc = rr if forward else nr - 1 - rr
                T = tab[fr + rloc]
                w = int(T[43])
                ext = 0
                for k in range(w & 7):
                    if (cta_index >> ((w >> (8 + 6 * k)) & 63)) & 1:
                        ext ^= int(T[44 + k])
                basew = T[tid & 15] ^ T[16 + (tid >> 4)] ^ np.uint32(ext)
                ro = [int(T[32 + a]) for a in range(r)]
                ms = [int(T[38 + a]) for a in range(r)]
                for a in range(r):
                    if ms[a] != P.LT_EMPTY:
                        g = gates[ms[a]]
                        fm = (0xFFFF if g["cf"] else 0) | (0xFFFF0000 if g["rf"] else 0)
                        basew = basew ^ np.uint32(ro[a] & fm)
                xm = np.zeros(NV, dtype=np.uint32)
                for a in range(r):
                    xm ^= np.where((jj >> a) & 1, np.uint32(ro[a]), np.uint32(0)).astype(np.uint32)
                full = basew[:, None] ^ xm[None, :]  # [NT, NV] packed load | store << 16
                la = (full & 0xFFFF).astype(np.int64)
                sa = (full >> 16).astype(np.int64)
                assert np.all(la % ebytes == 0) and np.all(sa % ebytes == 0)
                la //= ebytes
                sa //= ebytes
                assert np.array_equal(np.sort(la.reshape(-1)), e_idx), "load addresses must cover the tile exactly once"
                assert np.array_equal(np.sort(sa.reshape(-1)), e_idx), "store addresses must cover the tile exactly once"
                if not (int(T[37]) & 1):  # no barrier between loads and stores: every thread must store where it loaded
                    assert np.array_equal(np.sort(la, axis=1), np.sort(sa, axis=1))
                p = smem_p[la]
                l = smem_l[la] if lam is not None else None
                order = range(r) if forward else reversed(range(r))
                for a in range(r):  # adjoint: every moment of the round from the registers as loaded
                    if ms[a] == P.LT_EMPTY or not gates[ms[a]]["grad"]:
                        continue
                    g = gates[ms[a]]
                    lo = jj[(jj >> a) & 1 == 0]
                    hi = lo | (1 << a)
                    assert lam is not None and grads is not None
                    mo = _moments(p, l, lo, hi)
                    sg = -1.0 if g["cf"] else 1.0
                    m4 = {"I": g["mscale"] * mo[0], "X": g["mscale"] * mo[1], "Y": sg * g["mscale"] * mo[2], "Z": sg * g["mscale"] * mo[3]}
                    _chain_gradients(plan, g["recs"], m4, params, b, grads)
                for a in order:
                    if ms[a] == P.LT_EMPTY:
                        continue
                    g = gates[ms[a]]
                    lo = jj[(jj >> a) & 1 == 0]
                    hi = lo | (1 << a)
                    for arr in ([p] if lam is None else [p, l]):
                        x, y = arr[:, lo].copy(), arr[:, hi].copy()
                        arr[:, lo] = x + g["beta"] * y
                        arr[:, hi] = g["gamma"] * x + g["delta"] * y
                if rr == nr - 1:
                    p = p * sscal
                    if lam is not None:
                        l = l * sscal
                new_p = np.zeros_like(smem_p)
                new_p[sa] = p
                smem_p = new_p
                if lam is not None:
                    new_l = np.zeros_like(smem_l)
                    new_l[sa] = l
                    smem_l = new_l
            psi[b, base + dep] = smem_p[lay]
            if lam is not None:
                lam[b, base + dep] = smem_l[lay]


def _check_tma_coords(tm, c: int, b: int, base: int, n: int, tile_bits) -> None:
    """The box origin the kernel computes from (tile counter, batch element) must be the tile's base offset, and the box
    must enumerate exactly the tile bits in ascending order."""
    rank = int(tm["rank"]) & 0xFF
    off = 0
    enum_bits = []
    for d in range(rank):
        st, nb, it, src = int(tm["start"][d]), int(tm["nbits"][d]), int(tm["is_tile"][d]), int(tm["src"][d])
        if it:
            enum_bits += list(range(st, st + nb))
            crd = 0
        else:
            crd = (c >> src) & ((1 << nb) - 1)
        if d == rank - 1:
            assert not it and st + nb == n
            crd += b << nb
        off += crd << st
    assert enum_bits == sorted(tile_bits), (enum_bits, tile_bits)
    assert off == base + (b << n), (off, base, b)


def _moments(p, l, lo, hi):
    p0, p1, l0, l1 = p[:, lo], p[:, hi], l[:, lo], l[:, hi]
    s00 = np.sum(np.imag(np.conj(l0) * p0))
    s11 = np.sum(np.imag(np.conj(l1) * p1))
    x = np.sum(np.imag(np.conj(l0) * p1 + np.conj(l1) * p0))
    y = np.sum(np.real(np.conj(l1) * p0 - np.conj(l0) * p1))
    return s00 + s11, x, y, s00 - s11


def _chain_gradients(plan, recs, mom, params, b, grads) -> None:
    mats = [_gate_matrix(plan, c, params, b, False)[1] for c in recs]
    W = np.eye(2, dtype=np.complex128)
    for i in range(len(recs) - 1, -1, -1):
        c = recs[i]
        if int(c["flags"]) & P.GF_GRAD:
            kind = int(c["kind"])
            Pm = {P.GK_RX: _SIG["X"], P.GK_RY: _SIG["Y"], P.GK_RZ: _SIG["Z"], P.GK_PHASE: _SIG["Z"]}[kind]
            cI = -1.0 if kind == P.GK_PHASE else 0.0
            A = W @ Pm @ W.conj().T
            az, ax, ay = A[0, 0].real, A[1, 0].real, A[1, 0].imag
            grads[int(c["grad_row"]), b] += cI * mom["I"] + ax * mom["X"] + ay * mom["Y"] + az * mom["Z"]
        W = W @ mats[i]


def emulate_lean(plan: P.SweepPlan, psi: np.ndarray, params: np.ndarray, lam: np.ndarray | None = None):
    """Forward (lam is None) or adjoint execution of a plan whose sweeps are ALL lean.  Returns (psi, lam, grads)."""
    psi = psi.astype(np.complex128).copy()
    forward = lam is None
    lam = None if lam is None else lam.astype(np.complex128).copy()
    grads = np.zeros_like(params, dtype=np.float64)
    sweeps = range(plan.n_sweeps) if forward else reversed(range(plan.n_sweeps))
    for si in sweeps:
        emulate_lean_sweep(plan, si, psi, lam, params, forward, grads)
    return psi, lam, grads


def check_tables_hazard_free(plan: P.SweepPlan, forward: bool) -> int:
    """Static race check of the lean tables (the substitute for `compute-sanitizer --tool racecheck`, which is closed on
    this GPU pool): for EVERY round of every lean sweep, with every combination of pivot flips and external-control
    constants, (1) the threads' load addresses cover the tile exactly once, (2) so do the store addresses — no two
    threads ever write one shared-memory word —, and (3) a round without the barrier between its loads and stores has every
    thread store exactly where it loaded (nobody reads what another thread overwrites).  Returns the number of rounds."""
    tab = plan.lean_tab_fwd if forward else plan.lean_tab_bwd
    c128 = plan.cfg.c128
    ebytes = 16 if c128 else 8
    n_checked = 0
    for sw in plan.sweeps:
        if int(sw["io_mode"]) == P.IO_GENERIC:
            continue
        m, r = int(sw["m"]), int(sw["r"])
        NT, NV, TILE = 1 << (m - r), 1 << r, 1 << m
        tid, jj, e_idx = np.arange(NT), np.arange(NV), np.arange(TILE)
        for k in range(int(sw["n_rounds"])):
            T = tab[int(sw["first_round"]) + k]
            ro = [int(T[32 + a]) for a in range(r)]
            npx = int(T[43]) & 7
            exts = [0]
            for q in range(npx):
                exts = exts + [x ^ int(T[44 + q]) for x in exts]
            xm = np.zeros(NV, dtype=np.uint32)
            for a in range(r):
                xm ^= np.where((jj >> a) & 1, np.uint32(ro[a]), np.uint32(0)).astype(np.uint32)
            for ext in exts:
                for flips in range(1 << (2 * r)):  # column / row flip of every slot
                    fx = 0
                    for a in range(r):
                        fx ^= ro[a] & ((0xFFFF if (flips >> (2 * a)) & 1 else 0) | (0xFFFF0000 if (flips >> (2 * a + 1)) & 1 else 0))
                    basew = T[tid & 15] ^ T[16 + (tid >> 4)] ^ np.uint32(ext) ^ np.uint32(fx)
                    full = basew[:, None] ^ xm[None, :]
                    la, sa = (full & 0xFFFF).astype(np.int64), (full >> 16).astype(np.int64)
                    assert np.all(la % ebytes == 0) and np.all(sa % ebytes == 0)
                    assert np.array_equal(np.sort((la // ebytes).reshape(-1)), e_idx), "loads must cover the tile exactly once"
                    assert np.array_equal(np.sort((sa // ebytes).reshape(-1)), e_idx), "stores must cover the tile exactly once"
                    if not (int(T[37]) & 1):
                        assert np.array_equal(np.sort(la, axis=1), np.sort(sa, axis=1)), "no barrier: a thread must store where it loaded"
                    if r > 3 and flips > 16:  # the flip combinations are a group action on the register index: sample beyond 3 slots
                        break
            n_checked += 1
    return n_checked
