"""Fusion planner: a run of simple gates → fused sweeps → register rounds (host side, convert time).

Where the reference converts a block tree into one pyqtorch module per gate
(`/root/reference/qadence/backends/pyqtorch/convert_ops.py:273-345`, `pyq.Merge` only for chains on
a single qubit, `:277-278`), this planner groups gates so that each group costs ONE read + ONE write
of the state (a "sweep", `include/b200sv.h: b200sv_sweep`):

* a sweep stages tiles of 2^m amplitudes in shared memory; a non-diagonal gate fits if its target
  bit is one of the m tile bits (controls and diagonal gates fit anywhere: they are predicates /
  phases evaluated from the basis index);
* inside a sweep, gates are packed into rounds of r register bits;
* gates are pulled forward past deferred gates only when they commute with all of them
  (disjoint supports, or both acting diagonally — control or diagonal target — on every shared bit).

Pure Python/numpy; no CUDA needed (unit-tested on CPU against the oracle).
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Sequence

import numpy as np

from .program import Op, OpKind

# ---- binary layouts (must match include/b200sv.h) ---------------------------------------------
GATE_DT = np.dtype(
    [
        ("kind", "<i4"), ("a", "<i4"), ("a2", "<i4"), ("ext_bit", "<i4"), ("ctrl_ext", "<u8"),
        ("ctrl_reg", "<i4"), ("slot", "<i4"), ("cidx", "<i4"), ("grad_row", "<i4"), ("flags", "<i4"),
        ("chain", "<i4"), ("xidx", "<i4"), ("xkind", "<i4"),
    ]
)
PEXT_DT = np.dtype([("bit", "u1"), ("pad", "u1"), ("vec", "<u2")])
ROUND_DT = np.dtype(
    [
        ("first_gate", "<i4"), ("n_gates", "<i4"), ("regpos", "u1", (4,)), ("thrpos", "u1", (12,)), ("first_exec", "<i4"),
        ("n_exec", "<i4"), ("mslot", "<i4", (4,)), ("pcol", "<u2", (12,)), ("pconst", "<u2"), ("n_pext", "<u2"),
        ("pext", PEXT_DT, (4,)), ("flags", "<i4"), ("mslot4", "<i4"), ("regpos4", "u1"), ("pad4", "u1", (11,)),
    ]
)
SWEEP_DT = np.dtype(
    [
        ("m", "<i4"), ("r", "<i4"), ("first_round", "<i4"), ("n_rounds", "<i4"), ("first_gate", "<i4"),
        ("n_gates", "<i4"), ("n_outer", "<i4"), ("n_exec", "<i4"), ("tile_bits", "u1", (16,)), ("outer_bits", "u1", (44,)),
        ("io_mode", "u1"), ("pad_io", "u1", (3,)),
        ("first_exec", "<i4"), ("n_seg", "<i2"), ("n_gen", "<i2"), ("seg", np.dtype([("src", "u1"), ("len", "u1"), ("dst", "u1"), ("pad", "u1")]), (6,)),
    ]
)
PTERM_DT = np.dtype([("xmask", "<u8"), ("zmask", "<u8"), ("coef_row", "<i4"), ("n_y", "<i4"), ("pad", "<i4", (2,))])
TMA_DT = np.dtype([("rank", "<i4"), ("start", "u1", (5,)), ("nbits", "u1", (5,)), ("is_tile", "u1", (5,)), ("src", "u1", (5,))])
assert GATE_DT.itemsize == 56 and ROUND_DT.itemsize == 112 and SWEEP_DT.itemsize == 128 and PTERM_DT.itemsize == 32
assert TMA_DT.itemsize == 24
IO_GENERIC, IO_LEAN, IO_LEAN_TMA = 0, 1, 2
LT_STRIDE = 48
LT_EMPTY = 0xFFFF
TMA_SWIZZLE_FLAG = 0x100  # b200sv_tma.rank bit 8: the tensor map uses CU_TENSOR_MAP_SWIZZLE_128B
TMA_SWIZZLE_DEFAULT = "0"
VPERM_DEFAULT = "1"  # track the folded permutations of a lean sweep instead of applying them round by round (_lean_sweep_tables)

GK_MAT, GK_REAL, GK_RX, GK_RY, GK_RZ, GK_PHASE, GK_DIAG, GK_X, GK_SWAP, GK_SCALE = range(10)
GF_GRAD = 1
GF_MSLOT = 2
RF_PERM = 1
XK_PERM = 7
MAX_PEXT = 4
XK_MAT, XK_REAL, XK_RX, XK_DIAG, XK_X, XK_SWAP, XK_SCALE = range(7)
_XK_SINGLE = {GK_MAT: XK_MAT, GK_REAL: XK_REAL, GK_RX: XK_RX, GK_RY: XK_REAL, GK_RZ: XK_DIAG, GK_PHASE: XK_DIAG,
              GK_DIAG: XK_DIAG, GK_X: XK_X, GK_SWAP: XK_SWAP, GK_SCALE: XK_SCALE}

MAX_GATES_PER_SWEEP = 96
_X = np.array([[0, 1], [1, 0]], dtype=np.complex128)


@dataclass
class TileConfig:
    """Tile geometry for one (dtype, mode): m tile bits, r register bits, `low` forced low bits."""

    m: int
    r: int
    low: int
    fold: int  # bank-swizzle fold width: 3 for 16-byte amplitudes, 4 for 8-byte
    adjoint: bool = False  # direction the plan will be executed in (decides which kernels exist for it)
    lean: bool = True  # let sweeps without generic-phase entries run on lean_sweep_kernel (B200SV_LEAN=0 disables)

    @property
    def c128(self) -> bool:
        return self.fold == 3

    @staticmethod
    def default(dtype_is_c128: bool, adjoint: bool) -> "TileConfig":
        import os

        key = f"B200SV_TILE_{'ADJ' if adjoint else 'FWD'}_{'C128' if dtype_is_c128 else 'C64'}"
        if key in os.environ:  # tuning override: "m,r,low"
            m, r, low = (int(x) for x in os.environ[key].split(","))
            return TileConfig(m=m, r=r, low=low, fold=3 if dtype_is_c128 else 4, adjoint=adjoint)
        # measured on B200 (profiles/r01_tile_geometry.md): 64-byte runs (low = 2 for 16-byte amplitudes) already
        # stream at ~80 % of the HBM roofline, and every forced low bit costs a "hard" bit, i.e. more sweeps
        if dtype_is_c128:
            # forward: 2^10-amplitude tiles with 64-thread CTAs (7 per SM) beat 2^11 / 128 threads (3 per SM, the tile
            # alignment slack costs the fourth) although cfg2 then needs 9 sweeps instead of 8: 9 x 3.17 ms vs 8 x 3.87 ms
            # (profiles/r02_lean_kernel.md)
            # adjoint: r = 4 (16 amplitudes of psi and lambda per thread, 255 registers, 4 CTAs of 64 threads per SM): 48 instead
            # of 61 rounds for cfg2 — the per-round overhead (tile exchange, moment reduction) outweighs the 8 instead of 16
            # warps per SM: 85.6 ms per pass (9 sweeps) vs 93.2 ms with r = 3 (8 sweeps)
            return TileConfig(m=10, r=4, low=2, fold=3) if not adjoint else TileConfig(m=10, r=4, low=2, fold=3, adjoint=True)
        # complex64: the sweeps are issue bound (FFMA is not the limit), so the per-round overhead decides: more register
        # bits per round.  cfg2, ms per pass (profiles/r02_lean_kernel.md §7): forward (12,4) 22.9 -> (12,5) 18.1; adjoint
        # (11,3) 70.8 -> (11,4) 52.1
        return TileConfig(m=12, r=5, low=2, fold=4) if not adjoint else TileConfig(m=11, r=4, low=2, fold=4, adjoint=True)


@dataclass
class _G:
    """A simple gate in bit-position space."""

    kind: int
    hard: tuple[int, ...]  # bit positions that must be register bits (non-diagonal targets)
    soft: int  # bit position of a diagonal target, or -1
    ctrl: tuple[int, ...]
    slot: int
    cidx: int
    grad: bool
    src: int  # index in the op list (diagnostics)

    @property
    def bits_any(self) -> set[int]:
        s = set(self.hard) | set(self.ctrl)
        if self.soft >= 0:
            s.add(self.soft)
        return s


class ConstPool:
    def __init__(self) -> None:
        self.data: list[complex] = []
        self._index: dict[tuple, int] = {}

    def add(self, vals: Sequence[complex]) -> int:
        key = tuple(complex(v) for v in vals)
        if key not in self._index:
            self._index[key] = len(self.data)
            self.data.extend(key)
        return self._index[key]

    def array(self) -> np.ndarray:
        a = np.zeros((max(1, len(self.data)), 2), dtype=np.float64)
        for i, v in enumerate(self.data):
            a[i, 0], a[i, 1] = v.real, v.imag
        return a


def _lower(ops: Sequence[Op], n: int, slot_of: dict[str, int], pool: ConstPool, r: int = 2, harden_diag: bool = False) -> list[_G]:
    """`n` is the TOTAL qubit count: bit positions >= n_local refer to the rank bits of a sharded state.
    `harden_diag`: uncontrolled diagonal 1-qubit gates ask for their target as a tile bit like any other 1-qubit gate, so
    that they merge into the 2x2 chains of lean sweeps instead of becoming interpreted "external diagonal" entries."""
    out: list[_G] = []
    for i, op in enumerate(ops):
        bit = lambda q: n - 1 - q  # noqa: E731
        ctrl = tuple(sorted(bit(q) for q in op.controls))
        slot, cidx, grad = -1, -1, False
        if op.kind in (OpKind.RX, OpKind.RY, OpKind.RZ, OpKind.PHASE, OpKind.SCALE):
            if isinstance(op.param, str):
                slot, grad = slot_of[op.param], True
            else:
                cidx = pool.add([complex(op.param)])
        if op.kind == OpKind.MAT1:
            m = np.asarray(op.matrix, dtype=np.complex128).reshape(2, 2)
            t = bit(op.targets[0])
            if m[0, 1] == 0 and m[1, 0] == 0:
                if harden_diag and not ctrl:
                    out.append(_G(GK_DIAG, (t,), -1, ctrl, -1, pool.add([m[0, 0], m[1, 1]]), False, i))
                else:
                    out.append(_G(GK_DIAG, (), t, ctrl, -1, pool.add([m[0, 0], m[1, 1]]), False, i))
            elif np.array_equal(m, _X):
                out.append(_G(GK_X, (t,), -1, ctrl, -1, -1, False, i))
            elif np.all(m.imag == 0):
                out.append(_G(GK_REAL, (t,), -1, ctrl, -1, pool.add([complex(m[0, 0].real, m[0, 1].real), complex(m[1, 0].real, m[1, 1].real)]), False, i))
            else:
                out.append(_G(GK_MAT, (t,), -1, ctrl, -1, pool.add(list(m.ravel())), False, i))
        elif op.kind in (OpKind.RX, OpKind.RY):
            out.append(_G(GK_RX if op.kind == OpKind.RX else GK_RY, (bit(op.targets[0]),), -1, ctrl, slot, cidx, grad, i))
        elif op.kind in (OpKind.RZ, OpKind.PHASE):
            if harden_diag and not ctrl:
                out.append(_G(GK_RZ if op.kind == OpKind.RZ else GK_PHASE, (bit(op.targets[0]),), -1, ctrl, slot, cidx, grad, i))
            else:
                out.append(_G(GK_RZ if op.kind == OpKind.RZ else GK_PHASE, (), bit(op.targets[0]), ctrl, slot, cidx, grad, i))
        elif op.kind == OpKind.SWAP:
            ta, tb = (bit(q) for q in op.targets)
            if r >= 2:
                out.append(_G(GK_SWAP, tuple(sorted((ta, tb))), -1, ctrl, -1, -1, False, i))
            else:  # a single register bit cannot hold both targets: SWAP = 3 CNOTs (block_to_tensor.py:207-223)
                for c, t in ((ta, tb), (tb, ta), (ta, tb)):
                    out.append(_G(GK_X, (t,), -1, tuple(sorted((*ctrl, c))), -1, -1, False, i))
        elif op.kind == OpKind.SCALE:
            out.append(_G(GK_SCALE, (), -1, (), slot, cidx, grad, i))
        else:
            raise ValueError(f"op kind {op.kind} is not a simple gate")
    return out


def _pack(gates: list[_G], capacity: int, forced: set[int], max_gates: int) -> list[tuple[list[_G], set[int]]]:
    """Greedy commutation-aware packing of `gates` (ordered) into groups whose hard bits ∪ forced fit capacity."""
    groups: list[tuple[list[_G], set[int]]] = []
    pending = list(gates)
    while pending:
        bits = set(forced)
        taken: list[_G] = []
        deferred: list[_G] = []
        def_any: set[int] = set()  # every bit touched by a deferred gate
        def_hard: set[int] = set()  # bits a deferred gate acts on non-diagonally
        for g in pending:
            hard = set(g.hard)
            commutes = not (g.bits_any & def_hard) and not (hard & def_any)
            fits = len(bits | hard) <= capacity and len(taken) < max_gates
            if commutes and fits:
                taken.append(g)
                bits |= hard
            else:
                deferred.append(g)
                def_any |= g.bits_any
                def_hard |= hard
        if not taken:  # cannot happen: the first pending gate always commutes with the empty set
            raise RuntimeError("planner stalled (gate needs more register bits than the tile offers)")
        groups.append((taken, bits))
        pending = deferred
    return groups


# ---- tile search ------------------------------------------------------------------------------------
# `_pack` takes the tile bits in the order the gates ask for them.  On layered circuits that leaves short tail sweeps and
# half-empty tiles (hea(20, 8): 11 forward / 13 adjoint sweeps).  `_pack_search` chooses each sweep's tile among a few
# candidates — first-fit (with the first k frontier gates postponed) and contiguous windows of index bits, which let a sweep
# run a whole light-cone "trapezoid" of several layers — by a small beam search on the cost model
#   cost(plan) = Σ_sweeps (STREAM_COST + n_rounds)
# (one streaming pass of the state costs about three rounds of in-tile work, profiles/r01h_final_build.md).  The acceptance
# rule is `_pack`'s — a gate joins a sweep only if it commutes with every gate deferred before it — so the execution stays
# equivalent to the source order by the same argument; hea(20, 8) becomes 8 forward / 8 adjoint sweeps.
STREAM_COST = 3.0
SEARCH_WINDOW = 6 * MAX_GATES_PER_SWEEP  # pending gates looked at per sweep (keeps planning linear in the gate count)
BEAM_WIDTH = 6  # 3 left cfg2 with a permutation-only ninth sweep and 48 rounds; from 5 on: 8 sweeps, 44 rounds (planning 0.2 - 0.3 s)


def _take(pending: list[_G], tile: set[int], max_gates: int) -> tuple[list[_G], list[_G]]:
    """Gates of `pending` (ordered) a sweep over `tile` executes, and the rest (still ordered)."""
    taken: list[_G] = []
    deferred: list[_G] = []
    def_any: set[int] = set()
    def_hard: set[int] = set()
    head, tail = pending[:SEARCH_WINDOW], pending[SEARCH_WINDOW:]
    for g in head:
        hard = set(g.hard)
        if hard <= tile and len(taken) < max_gates and not (g.bits_any & def_hard) and not (hard & def_any):
            taken.append(g)
        else:
            deferred.append(g)
            def_any |= g.bits_any
            def_hard |= hard
    return taken, deferred + tail


def _firstfit_bits(pending: list[_G], capacity: int, forced: set[int], skip: int) -> set[int]:
    """Tile bits `_pack` would choose, after postponing the first `skip` frontier gates that ask for new bits."""
    bits = set(forced)
    def_any: set[int] = set()
    def_hard: set[int] = set()
    for g in pending[:SEARCH_WINDOW]:
        hard = set(g.hard)
        commutes = not (g.bits_any & def_hard) and not (hard & def_any)
        wants_new = bool(hard) and not hard <= bits
        if commutes and len(bits | hard) <= capacity and not (skip > 0 and wants_new):
            bits |= hard
        else:
            if commutes and skip > 0 and wants_new:
                skip -= 1
            def_any |= g.bits_any
            def_hard |= hard
    return bits


def _candidate_tiles(pending: list[_G], n: int, m: int, forced: set[int]) -> list[frozenset]:
    cands: list[set[int]] = [_firstfit_bits(pending, m, forced, skip) for skip in range(6)]
    w = m - len(forced)
    for s in range(len(forced), n - w + 1):
        cands.append(set(forced) | set(range(s, s + w)))
    out: list[frozenset] = []
    for t in cands:
        for p in range(n):  # complete with the lowest unused bit positions (as build_plan does)
            if len(t) >= m:
                break
            t.add(p)
        ft = frozenset(t)
        if ft not in out:
            out.append(ft)
    return out


def _pack_search(gates: list[_G], n: int, m: int, r: int, forced: set[int], max_gates: int, chains: bool) -> list[tuple[list[_G], set[int]]]:
    if n <= m or not gates:
        return _pack(gates, m, forced, max_gates)
    states: list[tuple[float, list[_G], list]] = [(0.0, list(gates), [])]
    done: list[tuple[float, list[_G], list]] = []
    while states:
        nxt: list[tuple[float, list[_G], list]] = []
        for cost, pending, groups in states:
            opts = []
            for tile in _candidate_tiles(pending, n, m, forced):
                taken, deferred = _take(pending, set(tile), max_gates)
                if not taken:
                    continue
                try:
                    n_rounds = len(_pack_rounds(taken, r, set(tile), chains))
                except RuntimeError:
                    continue
                work = sum(1.0 if not _is_perm(g) else 0.05 for g in taken)
                opts.append((work / (STREAM_COST + n_rounds), taken, deferred, tile, n_rounds))
            if not opts:
                raise RuntimeError("planner stalled (gate needs more register bits than the tile offers)")
            opts.sort(key=lambda o: -o[0])
            for _, taken, deferred, tile, n_rounds in opts[:BEAM_WIDTH]:
                st = (cost + STREAM_COST + n_rounds, deferred, groups + [(taken, set(tile))])
                (nxt if deferred else done).append(st)
        # rank partial plans by cost so far + an optimistic estimate of the rest (full rounds, no further streaming)
        nxt.sort(key=lambda st: st[0] + sum(1 for g in st[1] if not _is_perm(g)) / r)
        states = nxt[:BEAM_WIDTH]
    best = min(done, key=lambda st: st[0])
    # the beam is a heuristic: never return something the first-fit packer does better
    greedy = _pack(gates, m, forced, max_gates)
    greedy_cost = 0.0
    for sgates, bits in greedy:
        tile = set(bits)
        for p in range(n):
            if len(tile) >= m:
                break
            tile.add(p)
        greedy_cost += STREAM_COST + len(_pack_rounds(sgates, r, tile, chains))
    return best[2] if best[0] < greedy_cost else greedy


@dataclass
class _Round:
    """One round of a sweep: M chains (global target bit -> gates), G gates, P gates."""

    mchains: dict  # target bit -> list[_G]
    ggates: list  # list[list[_G]] (chains of length >= 1, executed by the interpreter)
    pgates: list  # list[_G]
    regbits: set


def _is_perm(g: _G) -> bool:
    return (g.kind == GK_X and len(g.ctrl) <= 1) or (g.kind == GK_SWAP and not g.ctrl)


def _is_matrix(g: _G) -> bool:
    return g.kind not in (GK_SWAP, GK_SCALE) and not g.ctrl


def _pack_rounds(sgates: list[_G], r: int, tile: set[int], phases: bool) -> list[_Round]:
    """Pack a sweep's gates (ordered) into rounds of the M / G / P shape (see b200sv_round).

    Acceptance keeps execution equivalent to the original order: a gate is taken only if it commutes
    with every deferred gate, and it may be placed in an earlier phase of the round only if it
    commutes with everything already accepted into the later phases.
    """
    rounds: list[_Round] = []
    pending = list(sgates)
    while pending:
        rd = _Round({}, [], [], set())
        deferred: list[_G] = []
        def_any: set[int] = set()
        def_hard: set[int] = set()
        g_any: set[int] = set()  # bits touched by accepted G gates
        g_hard: set[int] = set()
        p_any: set[int] = set()  # ... by accepted P gates
        p_hard: set[int] = set()
        open_g: dict[int, int] = {}  # target bit -> index of an open chain among G entries
        n_pext = set()
        n_taken = 0
        for g in pending:
            hard = set(g.hard)
            ok = not (g.bits_any & def_hard) and not (hard & def_any) and n_taken < MAX_GATES_PER_SWEEP
            placed = False
            if ok and phases and _is_perm(g) and all(b in tile for b in g.hard):
                ext = {c for c in g.ctrl if c not in tile}
                if len(n_pext | ext) <= MAX_PEXT:
                    rd.pgates.append(g)
                    n_pext |= ext
                    p_any |= g.bits_any
                    p_hard |= hard
                    placed = True
            elif ok and phases and _is_matrix(g):
                t = g.hard[0] if g.hard else g.soft
                before_gp = not (g.bits_any & (g_hard | p_hard)) and not (hard & (g_any | p_any))
                if before_gp and t in rd.mchains:
                    rd.mchains[t].append(g)
                    placed = True
                elif before_gp and t in tile and len(rd.regbits | {t}) <= r:
                    rd.mchains[t] = [g]
                    rd.regbits.add(t)
                    placed = True
            # an uncontrolled 1-qubit gate on a tile bit that cannot join the M phase of THIS round (it follows a permutation
            # or a generic gate on its bit) waits for the next round instead of entering the interpreted G phase: rounds
            # stay "lean" (merged 2x2 + permutations only), which is what lean_sweep_kernel executes
            waits = phases and _is_matrix(g) and (g.hard[0] if g.hard else g.soft) in tile
            if ok and not placed and not waits and not (phases and _is_perm(g) and all(b in tile for b in g.hard)):
                # generic phase: must commute with the accepted permutation gates; needs registers for hard bits
                before_p = not (g.bits_any & p_hard) and not (hard & p_any)
                if before_p and len(rd.regbits | hard) <= r:
                    one_q = g.kind not in (GK_SWAP, GK_SCALE)
                    t = (g.hard[0] if g.hard else g.soft) if one_q else -1
                    if one_q and not g.ctrl and t in open_g:
                        rd.ggates[open_g[t]].append(g)
                    else:
                        for bit in g.bits_any:
                            open_g.pop(bit, None)
                        rd.ggates.append([g])
                        if one_q and not g.ctrl:
                            open_g[t] = len(rd.ggates) - 1
                    rd.regbits |= hard
                    g_any |= g.bits_any
                    g_hard |= hard
                    placed = True
            if placed:
                n_taken += 1
            else:
                deferred.append(g)
                def_any |= g.bits_any
                def_hard |= hard
        if n_taken == 0:
            raise RuntimeError("planner stalled while packing rounds")
        rounds.append(rd)
        pending = deferred
    return rounds


def _perm_map(pgates: list[_G], local: dict[int, int]) -> tuple[list[int], int, dict[int, int]]:
    """Compose X / CNOT / SWAP gates into pi(i) = XOR_{bits of i} col[bit] ^ const ^ XOR_ext vec."""
    m = len(local)
    cols = [1 << i for i in range(m)]
    const = 0
    ext: dict[int, int] = {}

    def each(fn):
        nonlocal const
        for i in range(m):
            cols[i] = fn(cols[i])
        const = fn(const)
        for k in list(ext):
            ext[k] = fn(ext[k])

    for g in pgates:
        if g.kind == GK_X and not g.ctrl:
            const ^= 1 << local[g.hard[0]]
        elif g.kind == GK_X:
            t = local[g.hard[0]]
            c = g.ctrl[0]
            if c in local:
                lc = local[c]
                each(lambda v: v ^ (((v >> lc) & 1) << t))
            else:
                ext[c] = ext.get(c, 0) ^ (1 << t)
        elif g.kind == GK_SWAP:
            a, b = local[g.hard[0]], local[g.hard[1]]

            def sw(v, a=a, b=b):
                x = ((v >> a) ^ (v >> b)) & 1
                return v ^ ((x << a) | (x << b))

            each(sw)
        else:
            raise ValueError("not a permutation gate")
    return cols, const, ext


def _chain_xkind(chain: list[_G]) -> int:
    if len(chain) == 1:
        return _XK_SINGLE[chain[0].kind]
    kinds = {g.kind for g in chain}
    if kinds <= {GK_RZ, GK_PHASE, GK_DIAG}:
        return XK_DIAG
    if kinds <= {GK_RY, GK_REAL, GK_X}:
        return XK_REAL
    if kinds == {GK_RX}:
        return XK_RX
    return XK_MAT


def _thread_positions(m: int, regpos: list[int], fold: int) -> list[int]:
    avail = [p for p in range(m) if p not in regpos]
    chosen: list[int] = []
    used_res: set[int] = set()
    for _ in range(min(fold, len(avail))):
        pick = next((p for p in avail if p % fold not in used_res), None)
        if pick is None:
            pick = avail[0]
        chosen.append(pick)
        used_res.add(pick % fold)
        avail.remove(pick)
    return chosen + avail


@dataclass
class SweepPlan:
    """Packed, device-uploadable plan for one run of simple gates."""

    n_qubits: int
    cfg: TileConfig
    sweeps: np.ndarray
    rounds: np.ndarray
    gates: np.ndarray
    consts: np.ndarray  # [n_const, 2] float64
    n_source_gates: int
    exec_leader: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=np.int32))  # gate-record index per resolved entry
    lean_tab_fwd: np.ndarray = field(default_factory=lambda: np.zeros((0, LT_STRIDE), dtype=np.uint32))
    lean_tab_bwd: np.ndarray = field(default_factory=lambda: np.zeros((0, LT_STRIDE), dtype=np.uint32))
    tma: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=TMA_DT))
    exec_mode: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=np.uint8))  # io_mode of the owning sweep, per resolved entry
    # diagnostics: for each sweep, list of rounds, each a list of source-op indices (execution order)
    order: list[list[list[int]]] = field(default_factory=list)

    @property
    def n_sweeps(self) -> int:
        return int(self.sweeps.shape[0])

    @property
    def n_exec_total(self) -> int:
        return int(self.exec_leader.shape[0])

    def execution_order(self) -> list[int]:
        return [i for sw in self.order for rd in sw for i in rd]


def build_plan(ops: Sequence[Op], n: int, slot_of: dict[str, int], cfg: TileConfig, chains: bool = True,
               n_total: int | None = None, packer: str | None = None) -> SweepPlan:
    """Plan a run of simple gates on `n` LOCAL qubits.  With `n_total > n` the state is sharded by its
    highest-order qubits: qubit q maps to bit n_total-1-q, bits >= n are rank bits and may only be
    touched diagonally / as controls (evaluated from `index_hi` by the kernels)."""
    n_total = n if n_total is None else n_total
    pool = ConstPool()
    m = min(cfg.m, n)
    r = max(1, min(cfg.r, m)) if n > 0 else 1
    if m - r > 8:
        m = r + 8
    lowered = _lower(ops, n_total, slot_of, pool, r, harden_diag=(n_total == n and cfg.lean))
    for g_ in lowered:
        if any(b >= n for b in g_.hard):
            raise ValueError("non-diagonal gate on a global (rank) qubit: make it local first (dist.plan_layout)")
    forced = set(range(min(cfg.low, m)))
    # tile choice: "search" (beam search on the sweep/round cost model) or "greedy" (first fit; the sharded paths keep it:
    # their NVLink traffic was measured with it, and sweeps whose tiles avoid the rank bits stay local)
    if packer is None:
        import os

        packer = os.environ.get("B200SV_PLANNER", "search")
    if packer == "search" and n_total == n:
        sweep_groups = _pack_search(lowered, n, m, r, forced, MAX_GATES_PER_SWEEP, chains)
    else:
        sweep_groups = _pack(lowered, m, forced, MAX_GATES_PER_SWEEP)

    sweeps = np.zeros(len(sweep_groups), dtype=SWEEP_DT)
    rounds_l: list[np.ndarray] = []
    gates_l: list[np.ndarray] = []
    order: list[list[list[int]]] = []
    n_gates_total = 0
    n_rounds_total = 0
    n_exec_plan = 0
    for si, (sgates, bits) in enumerate(sweep_groups):
        # complete the tile with the lowest unused bit positions
        tile = sorted(bits)
        for p in range(n):
            if len(tile) >= m:
                break
            if p not in bits:
                tile.append(p)
        tile = sorted(tile)
        assert len(tile) == m, (tile, m)
        local = {p: i for i, p in enumerate(tile)}
        outer = [p for p in range(n) if p not in local]
        sw = sweeps[si]
        sw["m"], sw["r"], sw["n_outer"] = m, r, n - m
        sw["first_round"], sw["first_gate"] = n_rounds_total, n_gates_total
        sw["tile_bits"][:m] = tile
        sw["outer_bits"][: len(outer)] = outer
        # runs of consecutive outer bit positions -> (src bit of the CTA index, length, dst bit)
        segs: list[list[int]] = []
        for k, p in enumerate(outer):
            if segs and segs[-1][2] + segs[-1][1] == p:
                segs[-1][1] += 1
            else:
                segs.append([k, 1, p])
        if len(segs) <= 6:
            sw["n_seg"] = len(segs)
            for k, (src, ln, dst) in enumerate(segs):
                sw["seg"][k]["src"], sw["seg"][k]["len"], sw["seg"][k]["dst"] = src, ln, dst
        else:
            sw["n_seg"] = -1
        round_list = _pack_rounds(sgates, r, set(tile), chains)
        sw_order: list[list[int]] = []
        n_exec_sweep = 0
        for rnd in round_list:
            regs = sorted(local[p] for p in rnd.regbits)
            for p in range(m):  # pad register set to exactly r positions (prefer high tile bits: keeps low bits on threads)
                if len(regs) >= r:
                    break
                cand = m - 1 - p
                if cand not in regs:
                    regs.append(cand)
            regs = sorted(regs)
            reg_index = {tile[lp]: a for a, lp in enumerate(regs)}  # global bit -> register index
            m_entries = [rnd.mchains[t] for t in sorted(rnd.mchains, key=lambda t: reg_index[t])]
            entries = m_entries + rnd.ggates
            ordered = [g for ch in entries for g in ch] + list(rnd.pgates)
            rd = np.zeros(1, dtype=ROUND_DT)
            rd["first_gate"], rd["n_gates"] = n_gates_total, len(ordered)
            rd["regpos"][0, : min(r, 4)] = regs[:4]
            if r > 4:
                rd["regpos4"] = regs[4]
            thr = _thread_positions(m, regs, cfg.fold)
            rd["thrpos"][0, : len(thr)] = thr
            rd["mslot"][0, :] = -1
            rd["mslot4"] = -1
            ga = np.zeros(len(ordered), dtype=GATE_DT)
            gi = 0
            for ei, ch in enumerate(entries):
                is_m = ei < len(m_entries)
                xk = _chain_xkind(ch)
                if is_m:
                    ra = reg_index[ch[0].hard[0] if ch[0].hard else ch[0].soft]
                    if ra < 4:
                        rd["mslot"][0, ra] = n_exec_sweep
                    else:
                        rd["mslot4"] = n_exec_sweep
                elif ei == len(m_entries):
                    rd["first_exec"] = n_exec_sweep
                for ci, g in enumerate(ch):
                    rec = ga[gi]
                    gi += 1
                    rec["kind"], rec["slot"], rec["cidx"] = g.kind, g.slot, g.cidx
                    rec["a"], rec["a2"], rec["ext_bit"] = -1, -1, 0
                    rec["grad_row"] = g.slot if g.grad else -1
                    rec["flags"] = (GF_GRAD if g.grad else 0) | (GF_MSLOT if is_m else 0)
                    rec["chain"] = len(ch) if ci == 0 else 0
                    rec["xidx"], rec["xkind"] = n_exec_sweep, xk
                    if g.kind == GK_SWAP:
                        rec["a"], rec["a2"] = reg_index[g.hard[0]], reg_index[g.hard[1]]
                    elif g.hard:
                        rec["a"] = reg_index[g.hard[0]]
                    elif g.soft >= 0:
                        if g.soft in reg_index:
                            rec["a"] = reg_index[g.soft]
                        else:
                            rec["ext_bit"] = g.soft
                    cext, creg = 0, 0
                    for cb in g.ctrl:
                        if cb in reg_index:
                            creg |= 1 << reg_index[cb]
                        else:
                            cext |= 1 << cb
                    rec["ctrl_ext"], rec["ctrl_reg"] = cext, creg
                n_exec_sweep += 1
            if not rnd.ggates:
                rd["first_exec"] = n_exec_sweep
            rd["n_exec"] = len(rnd.ggates)
            for g in rnd.pgates:  # kept as records for diagnostics / emulation; folded into the address map
                rec = ga[gi]
                gi += 1
                rec["kind"], rec["slot"], rec["cidx"], rec["grad_row"] = g.kind, -1, -1, -1
                rec["a"], rec["a2"] = local[g.hard[0]], (local[g.hard[1]] if len(g.hard) > 1 else -1)  # TILE-local bits
                rec["ext_bit"] = g.ctrl[0] if g.ctrl else -1  # GLOBAL control bit
                rec["chain"], rec["xidx"], rec["xkind"] = 0, -1, XK_PERM
            cols, const, ext = _perm_map(rnd.pgates, local)
            rd["pcol"][0, :m] = cols
            rd["pconst"] = const
            ext = {k: v for k, v in ext.items() if v}
            rd["n_pext"] = len(ext)
            for k, (bit, vec) in enumerate(sorted(ext.items())):
                rd["pext"][0, k]["bit"], rd["pext"][0, k]["vec"] = bit, vec
            if const or ext or cols != [1 << i for i in range(m)]:
                rd["flags"] = RF_PERM
            rounds_l.append(rd)
            gates_l.append(ga)
            sw_order.append([g.src for g in ordered])
            n_gates_total += len(ordered)
            n_rounds_total += 1
        sw["n_rounds"] = len(round_list)
        sw["n_gates"] = n_gates_total - int(sw["first_gate"])
        sw["n_exec"] = n_exec_sweep
        sw["n_gen"] = sum(len(rnd.ggates) for rnd in round_list)
        sw["first_exec"] = n_exec_plan
        n_exec_plan += n_exec_sweep
        order.append(sw_order)
    rounds = np.concatenate(rounds_l) if rounds_l else np.zeros(0, dtype=ROUND_DT)
    gates = np.concatenate(gates_l) if gates_l else np.zeros(0, dtype=GATE_DT)
    leaders = np.zeros(n_exec_plan, dtype=np.int32)
    for si in range(len(sweeps)):
        fg, ng, fx = int(sweeps[si]["first_gate"]), int(sweeps[si]["n_gates"]), int(sweeps[si]["first_exec"])
        for gi in range(fg, fg + ng):
            if gates[gi]["chain"] >= 1:
                leaders[fx + int(gates[gi]["xidx"])] = gi
    plan = SweepPlan(n, cfg, sweeps, rounds, gates, pool.array(), len(ops), leaders, order=order)
    finish_lean(plan, n_total)
    if r > 4 and any(int(sw["io_mode"]) == IO_GENERIC for sw in plan.sweeps):
        # the generic kernel holds at most 4 register bits: a plan with sweeps the lean kernel cannot take (small registers,
        # controlled gates, sharded states) is rebuilt with r = 4
        from dataclasses import replace

        return build_plan(ops, n, slot_of, replace(cfg, r=4), chains, n_total, packer)
    if cfg.adjoint and cfg.c128 and r == 4 and any(int(sw["io_mode"]) == IO_GENERIC for sw in plan.sweeps):
        # ... and its complex128 adjoint variant was tuned with r = 3 (r = 4 would be 128 data registers of psi and lambda
        # next to the interpreter's state): keep that geometry for plans that are not purely lean
        from dataclasses import replace

        return build_plan(ops, n, slot_of, replace(cfg, r=3), chains, n_total, packer)
    return plan


# ---- lean sweeps: kernel choice, TMA boxes, index tables --------------------------------------------------
def _lean_supported(c128: bool, m: int, r: int, adjoint: bool) -> bool:
    """Whether `lean_sweep_kernel` is instantiated for this geometry (asks the library; host-only call)."""
    from . import cabi

    try:
        return bool(cabi.load().b200sv_lean_supported(cabi.C128 if c128 else cabi.C64, m, r, int(adjoint)))
    except cabi.B200svError:
        return False


def tma_box(tile_bits: Sequence[int], n: int, c128: bool, swizzle: bool = False) -> np.ndarray | None:
    """The tile as a box of the [batch][2^n] state seen as a tensor of 8-byte elements (see `b200sv_tma`), or None when it
    needs more than 5 dimensions.  Runs of consecutive tile bits become tile dimensions (at most 256 elements each, so long
    runs are split), the gaps between them dimensions walked one element at a time by the tile counter."""
    tile = sorted(int(t) for t in tile_bits)
    if not tile or tile[0] != 0:
        return None
    if not c128 and (len(tile) < 2 or tile[1] != 1):
        return None  # the innermost box row must be at least 16 bytes
    runs: list[list[int]] = []
    for t in tile:
        if runs and runs[-1][0] + runs[-1][1] == t:
            runs[-1][1] += 1
        else:
            runs.append([t, 1])
    dims: list[tuple[int, int, int, int]] = []  # (start bit, n bits, is_tile, src bit of the tile counter)
    pos = src = 0
    for start, length in runs:
        if start > pos:
            dims.append((pos, start - pos, 0, src))
            src += start - pos
        st, rem = start, length
        while rem > 0:
            cap = 8 - (1 if (c128 and st == 0) else 0)
            if swizzle and st == 0:
                cap = 3 if c128 else 4  # SWIZZLE_128B: the innermost box row is at most 128 bytes
            k = min(rem, cap)
            dims.append((st, k, 1, 0))
            st += k
            rem -= k
        pos = start + length
    dims.append((pos, n - pos, 0, src))  # remaining high bits x batch
    if len(dims) > 5 or len(dims) < 2:
        return None
    rec = np.zeros(1, dtype=TMA_DT)
    rec["rank"] = len(dims)
    for d, (st, nb, it, sr) in enumerate(dims):
        rec["start"][0, d], rec["nbits"][0, d], rec["is_tile"][0, d], rec["src"][0, d] = st, nb, it, sr
    return rec


def _swz_a(e: int, c128: bool) -> int:
    """Layout a TMA load with CU_TENSOR_MAP_SWIZZLE_128B leaves in shared memory: dense box order, then the 16-byte chunk
    index (byte-address bits 4..6) XOR-ed with address bits 7..9 (measured: tools/ubench/tma_probe.cu).  In amplitude
    units: complex128 (16 B) bits 0..2 ^= bits 3..5; complex64 (8 B) bits 1..3 ^= bits 4..6."""
    return e ^ ((e >> 3) & 7) if c128 else e ^ (((e >> 4) & 7) << 1)


def _swz_b(e: int, c128: bool) -> int:
    """Bank swizzle of the in-tile layout between rounds (`swz<real>` in sweep.cuh); GF(2)-linear."""
    return e ^ (((e >> 3) ^ (e >> 6) ^ (e >> 9)) & 7) if c128 else e ^ (((e >> 4) ^ (e >> 8)) & 15)


class _Affine:
    """x -> XOR_{bits i of x} cols[i] ^ const ^ XOR_{external bit e set in the tile index} ext[e] on GF(2)^m."""

    def __init__(self, cols: list[int], const: int = 0, ext: dict[int, int] | None = None):
        self.cols, self.const, self.ext = list(cols), int(const), {k: v for k, v in (ext or {}).items() if v}

    @staticmethod
    def identity(m: int) -> "_Affine":
        return _Affine([1 << i for i in range(m)])

    @staticmethod
    def of_round(rd, m: int) -> "_Affine":
        if not (int(rd["flags"]) & RF_PERM):
            return _Affine.identity(m)
        ext = {int(rd["pext"][k]["bit"]): int(rd["pext"][k]["vec"]) for k in range(int(rd["n_pext"]))}
        return _Affine([int(x) for x in rd["pcol"][:m]], int(rd["pconst"]), ext)

    def lin(self, v: int) -> int:
        out, i = 0, 0
        while v:
            if v & 1:
                out ^= self.cols[i]
            v >>= 1
            i += 1
        return out

    def after(self, first: "_Affine") -> "_Affine":
        """self ∘ first."""
        ext = dict(self.ext)
        for e, v in first.ext.items():
            ext[e] = ext.get(e, 0) ^ self.lin(v)
        return _Affine([self.lin(c) for c in first.cols], self.const ^ self.lin(first.const), ext)

    def inverse(self) -> "_Affine":
        m = len(self.cols)
        inv = {self.lin(x): x for x in range(1 << m)}
        assert len(inv) == 1 << m, "not a permutation"
        lin_inv = _Affine([inv[1 << i] for i in range(m)])
        return _Affine(lin_inv.cols, lin_inv.lin(self.const), {e: lin_inv.lin(v) for e, v in self.ext.items()})

    @property
    def is_identity(self) -> bool:
        return not self.const and not self.ext and self.cols == [1 << i for i in range(len(self.cols))]


def _rank_gf2(vs: list[int]) -> int:
    basis: list[int] = []
    for v in vs:
        for b_ in basis:
            v = min(v, v ^ b_)
        if v:
            basis.append(v)
    return len(basis)


def _lean_table(rd, m: int, r: int, c128: bool, lmap: _Affine, smap: _Affine, load_plain_layout: bool, store_plain_layout: bool,
                tma_swizzle: bool = False, reorder_threads: bool = False) -> np.ndarray:
    """One round's B200SV_LT_STRIDE words for one direction (layout documented in include/b200sv.h): the amplitude whose
    position in the ROUND'S OWN frame is x (thread bits -> `thrpos`, register bits -> `regpos`) is loaded from tile position
    lmap(x) and stored to smap(x).  `*_plain_layout`: that side of the round faces the TMA-written / TMA-read tile (box order,
    hardware-swizzled if `tma_swizzle`); otherwise the bank-swizzled layout between rounds."""
    shift = 4 if c128 else 3
    regs = [int(x) for x in rd["regpos"][: min(r, 4)]] + ([int(rd["regpos4"])] if r > 4 else [])
    thr = [int(x) for x in rd["thrpos"][: m - r]]

    def lay(v: int, plain_layout: bool) -> int:
        if plain_layout:
            return (_swz_a(v, c128) if tma_swizzle else v) << shift
        return _swz_b(v, c128) << shift

    def word(lv: int, sv: int) -> int:
        lo, hi = lay(lv, load_plain_layout), lay(sv, store_plain_layout)
        assert lo < 65536 and hi < 65536
        return lo | (hi << 16)

    if reorder_threads:
        # which positions sit on the low lane bits decides the shared-memory bank conflicts: a quarter warp (complex128,
        # 16-byte accesses; half warp for complex64) is conflict-free iff its lanes' offsets differ in the bank bits, i.e.
        # iff the bank-bit projections of the low lane bits' address vectors are linearly independent — on both sides
        nb = 3 if c128 else 4
        bsh, bmask = (4, 7) if c128 else (3, 15)

        def banks(p: int) -> tuple[int, int]:
            w = word(lmap.cols[p], smap.cols[p])
            return ((w & 0xFFFF) >> bsh) & bmask, ((w >> 16) >> bsh) & bmask

        chosen: list[int] = []
        lo_set: list[int] = []
        hi_set: list[int] = []
        for want_both in (True, False):
            for p_ in thr:
                if len(chosen) >= nb:
                    break
                if p_ in chosen:
                    continue
                bl, bh = banks(p_)
                up_l = _rank_gf2(lo_set + [bl]) > _rank_gf2(lo_set)
                up_h = _rank_gf2(hi_set + [bh]) > _rank_gf2(hi_set)
                if (up_l and up_h) or (not want_both and (up_l or up_h)):
                    chosen.append(p_)
                    lo_set.append(bl)
                    hi_set.append(bh)
        thr = chosen + [p_ for p_ in thr if p_ not in chosen]

    T = np.zeros(LT_STRIDE, dtype=np.uint32)
    cw = word(lmap.const, smap.const)
    for half in range(2):
        for v in range(16):
            lv = sv = 0
            for k in range(4):
                tb = k + 4 * half
                if tb < len(thr) and (v >> k) & 1:
                    lv ^= lmap.cols[thr[tb]]
                    sv ^= smap.cols[thr[tb]]
            T[16 * half + v] = word(lv, sv) ^ (cw if half == 0 else 0)
    for a, pos in enumerate(regs):
        T[32 + a] = word(lmap.cols[pos], smap.cols[pos])
    ms = [int(x) for x in rd["mslot"]] + [int(rd["mslot4"])]
    for a in range(5):
        T[38 + a] = ms[a] if (a < r and ms[a] >= 0) else LT_EMPTY
    ext_bits = sorted(set(lmap.ext) | set(smap.ext))
    if len(ext_bits) > 4:
        raise _TooManyExternalBits()
    w = len(ext_bits)
    for k, bit in enumerate(ext_bits):
        w |= bit << (8 + 6 * k)
        T[44 + k] = word(lmap.ext.get(bit, 0), smap.ext.get(bit, 0))
    T[43] = w
    words = [int(x) for x in T[:37]] + [int(x) for x in T[44:48]]
    T[37] = 1 if any((x & 0xFFFF) != (x >> 16) for x in words) else 0  # some thread stores where another one loads: barrier
    return T


class _TooManyExternalBits(Exception):
    pass


def _lean_sweep_tables(rds, m: int, r: int, c128: bool, plain_io: bool, tma_swizzle: bool, virtual: bool) -> tuple[list, list]:
    """Tables of all rounds of one lean sweep, forward and backward direction.

    `virtual=False`: every round physically applies its permutation (the X / CNOT / SWAP gates folded after its 2x2 gates) when
    it stores: load at x, barrier, store at pi_k(x).  `virtual=True`: the permutations are only TRACKED — round k loads the
    amplitude of position x from S_k(x) and stores it back in place (no barrier between the loads and the stores), with
    S_0 = identity and S_{k+1} = S_k ∘ pi_k^{-1}; only the last round of the sweep stores to the real positions pi_k(x).
    Backward (rounds in reverse order, un-permute before un-applying): loads from B_k(pi_k(x)), B_last = identity,
    B_{k-1} = B_k ∘ pi_k, in place, and the final round (k = 0) stores to x."""
    nr = len(rds)
    pis = [_Affine.of_round(rd, m) for rd in rds]
    ident = _Affine.identity(m)
    fwd, bwd = [], []
    if not virtual:
        for k, rd in enumerate(rds):
            fwd.append(_lean_table(rd, m, r, c128, ident, pis[k], plain_io and k == 0, plain_io and k == nr - 1, tma_swizzle))
            bwd.append(_lean_table(rd, m, r, c128, pis[k], ident, plain_io and k == nr - 1, plain_io and k == 0, tma_swizzle))
        return fwd, bwd
    S = ident
    for k, rd in enumerate(rds):
        last = k == nr - 1
        fwd.append(_lean_table(rd, m, r, c128, S, pis[k] if last else S, plain_io and k == 0, plain_io and last, tma_swizzle, True))
        S = S.after(pis[k].inverse())
    Bm = ident
    for k in range(nr - 1, -1, -1):
        G = Bm.after(pis[k])
        bwd.append(_lean_table(rds[k], m, r, c128, G, ident if k == 0 else G, plain_io and k == nr - 1, plain_io and k == 0, tma_swizzle, True))
        Bm = G
    bwd.reverse()
    return fwd, bwd


def finish_lean(plan: "SweepPlan", n_total: int) -> None:
    """Choose the kernel of every sweep and build what the lean kernel needs (tables for both directions, TMA boxes)."""
    import os

    cfg = plan.cfg
    n = plan.n_qubits
    ns = plan.n_sweeps
    plan.tma = np.zeros(ns, dtype=TMA_DT)
    plan.lean_tab_fwd = np.zeros((max(1, len(plan.rounds)), LT_STRIDE), dtype=np.uint32)
    plan.lean_tab_bwd = np.zeros((max(1, len(plan.rounds)), LT_STRIDE), dtype=np.uint32)
    plan.exec_mode = np.zeros(max(1, plan.n_exec_total), dtype=np.uint8)
    allow = cfg.lean and os.environ.get("B200SV_LEAN", "1") != "0" and n_total == n
    allow_tma = os.environ.get("B200SV_TMA", "1") != "0"
    tma_swizzle = os.environ.get("B200SV_TMA_SWIZZLE", TMA_SWIZZLE_DEFAULT) != "0"
    virtual = os.environ.get("B200SV_VPERM", VPERM_DEFAULT) != "0"
    for si in range(ns):
        sw = plan.sweeps[si]
        m, r = int(sw["m"]), int(sw["r"])
        mode = IO_GENERIC
        if allow and int(sw["n_gen"]) == 0 and m == cfg.m and _lean_supported(cfg.c128, m, r, cfg.adjoint):
            mode = IO_LEAN
            box = tma_box(sw["tile_bits"][:m], n, cfg.c128, tma_swizzle) if allow_tma else None
            if box is not None:
                mode = IO_LEAN_TMA
                plan.tma[si] = box[0]
                if tma_swizzle:
                    plan.tma[si]["rank"] |= TMA_SWIZZLE_FLAG
        sw["io_mode"] = mode
        fx = int(sw["first_exec"])
        plan.exec_mode[fx : fx + int(sw["n_exec"])] = mode
        if mode == IO_GENERIC:
            continue
        fr, nr = int(sw["first_round"]), int(sw["n_rounds"])
        plain_io = mode == IO_LEAN_TMA  # TMA writes / reads the tile un-swizzled
        rds = [plan.rounds[fr + k] for k in range(nr)]
        try:
            fwd, bwd = _lean_sweep_tables(rds, m, r, cfg.c128, plain_io, plain_io and tma_swizzle, virtual)
        except _TooManyExternalBits:  # more than 4 external control bits accumulate over the sweep: permute physically
            fwd, bwd = _lean_sweep_tables(rds, m, r, cfg.c128, plain_io, plain_io and tma_swizzle, False)
        for k in range(nr):
            plan.lean_tab_fwd[fr + k] = fwd[k]
            plan.lean_tab_bwd[fr + k] = bwd[k]


# ---- Pauli sums -----------------------------------------------------------------------------------
def pack_pauli_terms(paulis_per_term: Sequence[Sequence[tuple[int, str]]], n: int) -> np.ndarray:
    """Pack Pauli strings (qubit, 'X'|'Y'|'Z') into b200sv_pterm records; coef_row = term index."""
    arr = np.zeros(len(paulis_per_term), dtype=PTERM_DT)
    for t, ps in enumerate(paulis_per_term):
        x = z = ny = 0
        for q, p in ps:
            bit = 1 << (n - 1 - q)
            if p in ("X", "Y"):
                x |= bit
            if p in ("Z", "Y"):
                z |= bit
            if p == "Y":
                ny += 1
        arr[t]["xmask"], arr[t]["zmask"], arr[t]["coef_row"], arr[t]["n_y"] = x, z, t, ny
    return arr
