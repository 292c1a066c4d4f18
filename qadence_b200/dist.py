"""Multi-GPU execution: one process per GPU, `torch.distributed` (NCCL over NVLink 5 / NVSwitch).

The reference only has data-parallel training (`qadence/ml_tools/train_utils/accelerator.py:296-305`,
DDP over gloo/nccl); the two sharding axes below are the ones SURVEY.md §8(e) derives from the path:

* **parameter / feature batches** shard trivially (every batch element is an independent state
  vector): `shard_batch` + one all-reduce(SUM) of the gradient vector of shared parameters (C1).
  No data-path collective.
* **one large register** is sharded by its highest-order qubits: rank = the top g = log2(P) index
  bits, each rank holds 2^(n-g) contiguous amplitudes.  Local gates, and any control / diagonal gate
  on a global qubit, are communication-free (the kernels evaluate them from `index_hi`).  A
  non-diagonal gate on a global qubit triggers a global↔local qubit exchange: because the g
  highest LOCAL bits index P contiguous chunks, the exchange is exactly one all-to-all of
  contiguous chunks (C2), done in place in bounded staging pieces so that a 128 GiB shard
  (36 qubits on 8 GPUs) never needs a second copy.  Local re-indexing before / after the exchange
  is done by in-place SWAP gates folded into the sweep kernels' permutation phase.
"""

from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass, field
from typing import Mapping, Sequence

import torch
import torch.distributed as dist
from torch import Tensor

from .program import Op, OpKind, PauliSum, Program

A2A_STAGING_BYTES = 1 << 30


# ---------------------------------------------------------------------------------------------------
# process-group plumbing
# ---------------------------------------------------------------------------------------------------
def init_from_env(backend: str | None = None) -> tuple[int, int]:
    """(rank, world) from torchrun's environment; initialises the default group if needed."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world


# ---------------------------------------------------------------------------------------------------
# batch sharding (configs[1]-style parameter batches, configs[3] training batches)
# ---------------------------------------------------------------------------------------------------
def shard_range(batch: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [lo, hi) slice of the batch owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(values: Mapping[str, Tensor], batch: int, rank: int, world: int) -> dict[str, Tensor]:
    """Slice every `[batch]`-shaped entry; `[1]`-shaped (shared) parameters are replicated."""
    lo, hi = shard_range(batch, rank, world)
    out = {}
    for k, v in values.items():
        if isinstance(v, Tensor) and v.dim() >= 1 and v.shape[0] == batch and batch > 1:
            out[k] = v[lo:hi]
        else:
            out[k] = v
    return out


def allreduce_shared_gradients(params: Sequence[Tensor], group=None) -> None:
    """C1: one all-reduce(SUM) of the flattened gradient vector of the shared parameters —
    replaces DDP's bucket all-reduce (`accelerator.py:296-305`); 576 doubles at configs[3]."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    # one fused multi-tensor copy back (hundreds of one-element gradients: a copy kernel each would cost more than the
    # all-reduce itself — measured 2.1 ms for 576 parameters, profiles/r02_multi_gpu.md)
    pieces = [p.view_as(g) for p, g in zip(flat.split([g.numel() for g in grads]), grads)]
    torch._foreach_copy_(grads, pieces)


def gather_batch(local: Tensor, batch: int, group=None) -> Tensor:
    """All-gather per-rank `[b_r, ...]` results into `[batch, ...]` (only when a caller needs it)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_range(batch, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((mx, *local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(outs, sizes)])


# ---------------------------------------------------------------------------------------------------
# global-qubit sharding: layout planning (pure host logic, unit-tested on CPU)
# ---------------------------------------------------------------------------------------------------
def hard_qubits(op: Op) -> tuple[int, ...]:
    """Qubits an op acts on NON-diagonally (they must be local when the op runs)."""
    if op.kind in (OpKind.RZ, OpKind.PHASE, OpKind.SCALE):
        return ()
    if op.kind == OpKind.MAT1:
        m = op.matrix
        return () if (m[0, 1] == 0 and m[1, 0] == 0) else tuple(op.targets)
    if op.kind in (OpKind.RX, OpKind.RY, OpKind.SWAP):
        return tuple(op.targets)
    raise NotImplementedError(f"{op.kind.name} is not supported on a qubit-sharded register yet")


@dataclass
class LayoutStep:
    kind: str  # "gates" | "a2a"
    ops: list[Op] = field(default_factory=list)  # "gates": ops relabelled to PHYSICAL qubit labels


@dataclass
class LayoutPlan:
    n_qubits: int
    g: int  # number of global (rank) qubits
    steps: list[LayoutStep]
    final_layout: list[int]  # final_layout[q] = physical label of logical qubit q (identity if restored)
    n_exchanges: int


def _relabel(op: Op, layout: Sequence[int]) -> Op:
    return Op(op.kind, tuple(layout[q] for q in op.targets), tuple(layout[q] for q in op.controls), op.param, op.matrix, None, None, op.label)


def plan_layout(ops: Sequence[Op], n: int, g: int, restore: bool = True) -> LayoutPlan:
    """Insert global↔local exchanges so that every non-diagonal target is local when its gate runs.

    Physical label k < g is a rank bit (label 0 = MSB of the rank); labels g..2g-1 are the g
    highest local bits.  An exchange ("a2a") swaps label k with label g+k for every k < g.  Victims
    (logical qubits sent to the rank bits) are the local qubits whose next non-diagonal use is
    farthest away (Belady); they are first moved to labels g..2g-1 by in-place local SWAP gates.
    """
    if g == 0:
        return LayoutPlan(n, 0, [LayoutStep("gates", list(ops))], list(range(n)), 0)
    if n < 2 * g:
        raise ValueError("a register sharded over 2^g ranks needs at least 2g qubits")
    layout = list(range(n))  # logical -> physical
    hard = [hard_qubits(op) for op in ops]
    hmax = max((len(h) for h in hard), default=0)
    if n - g - max(0, hmax - 1) < g:
        # an exchange always swaps all g rank bits with g local qubits that the current gate does not need
        raise ValueError(f"{n} qubits over 2^{g} ranks: gates with {hmax} non-diagonal targets need at least {2 * g + hmax - 1} qubits "
                         "for the all-to-all path (PeerShardedSimulator has no such limit)")
    INF = len(ops) + 1
    # next_use[i][q]: first j >= i with q in hard[j]
    nxt = [INF] * n
    next_use: list[list[int]] = [[]] * (len(ops) + 1)
    next_use[len(ops)] = list(nxt)
    for i in range(len(ops) - 1, -1, -1):
        for q in hard[i]:
            nxt[q] = i
        next_use[i] = list(nxt)

    steps: list[LayoutStep] = [LayoutStep("gates")]
    n_ex = 0

    def emit(op: Op) -> None:
        if steps[-1].kind != "gates":
            steps.append(LayoutStep("gates"))
        steps[-1].ops.append(op)

    def local_swap(la: int, lb: int) -> None:
        """Swap the contents of physical labels la, lb (both local) and update the layout."""
        if la == lb:
            return
        emit(Op(OpKind.SWAP, (la, lb), (), label="layout-swap"))
        qa, qb = layout.index(la), layout.index(lb)
        layout[qa], layout[qb] = lb, la

    def exchange(victims: Sequence[int]) -> None:
        nonlocal n_ex
        # move victims to labels g..2g-1 (any order), then swap label k <-> g+k
        top = list(range(g, 2 * g))
        free = [lab for lab in top if layout.index(lab) not in victims]
        for v in victims:
            if layout[v] not in top:
                local_swap(layout[v], free.pop())
        steps.append(LayoutStep("a2a"))
        n_ex += 1
        for q in range(n):
            if layout[q] < g:
                layout[q] += g
            elif layout[q] < 2 * g:
                layout[q] -= g

    for i, op in enumerate(ops):
        need = [q for q in hard[i] if layout[q] < g]
        if need:
            local_qs = [q for q in range(n) if layout[q] >= g and q not in hard[i]]
            local_qs.sort(key=lambda q: (-next_use[i][q], -q))
            exchange(local_qs[:g])
            assert all(layout[q] >= g for q in hard[i])
        emit(_relabel(op, layout))

    if restore and layout != list(range(n)):
        if any(layout[q] < g for q in range(g)):  # some of logical 0..g-1 sit on rank bits in the wrong place
            local_qs = [q for q in range(n) if layout[q] >= g and q >= g]
            exchange(local_qs[:g])
        if any(layout[q] != q for q in range(g)):
            for k in range(g):  # logical k -> label g+k, then one exchange puts it on rank bit k
                local_swap(layout[k], g + k)
            steps.append(LayoutStep("a2a"))
            n_ex += 1
            for q in range(n):
                if layout[q] < g:
                    layout[q] += g
                elif layout[q] < 2 * g:
                    layout[q] -= g
        for q in range(g, n):  # remaining local permutation, cycle by cycle
            if layout[q] != q:
                local_swap(layout[q], q)
        assert layout == list(range(n)), layout
    steps = [s for s in steps if s.kind == "a2a" or s.ops]
    return LayoutPlan(n, g, steps, list(layout), n_ex)


# ---------------------------------------------------------------------------------------------------
# C2: in-place chunked all-to-all of the g highest local bits with the rank bits
# ---------------------------------------------------------------------------------------------------
def exchange_global_local(state: Tensor, world: int, group=None, staging_bytes: int = A2A_STAGING_BYTES) -> int:
    """In place: the amplitude block with (rank=r, top-local=s) moves to (rank=s, top-local=r).

    `state` is this rank's `[2^n_local]` (or `[1, 2^n_local]`) contiguous shard.  Returns the number
    of bytes this rank sent over the fabric (S_local·(1 − 1/P))."""
    flat = state.view(-1)
    total = flat.numel()
    assert total % world == 0
    chunk = total // world  # contiguous block per destination (top-local index = destination rank)
    esz = flat.element_size()
    piece = max(1, min(chunk, staging_bytes // (esz * world)))
    send = torch.empty(world * piece, dtype=flat.dtype, device=flat.device)
    recv = torch.empty_like(send)
    view = flat.view(world, chunk)
    sent = 0
    for off in range(0, chunk, piece):
        w = min(piece, chunk - off)
        s = send[: world * w].view(world, w)
        r = recv[: world * w].view(world, w)
        s.copy_(view[:, off : off + w])
        dist.all_to_all_single(r.view(-1), s.view(-1), group=group)
        view[:, off : off + w].copy_(r)
        sent += (world - 1) * w * esz
    return sent


# ---------------------------------------------------------------------------------------------------
# qubit-sharded simulator
# ---------------------------------------------------------------------------------------------------
class QubitShardedSimulator:
    """Runs a Program of simple gates on an n-qubit register sharded over `world` = 2^g GPUs."""

    def __init__(self, program: Program, rank: int, world: int, device: torch.device | str | None = None, group=None):
        from . import engine

        if world & (world - 1):
            raise ValueError("world size must be a power of two")
        self.program = program
        self.n = program.n_qubits
        self.g = world.bit_length() - 1
        self.n_local = self.n - self.g
        self.rank, self.world, self.group = rank, world, group
        self.device = engine._dev(device)
        self.names = program.param_names
        self.slot_of = {nm: i for i, nm in enumerate(self.names)}
        self.layout = plan_layout(program.ops, self.n, self.g, restore=True)
        self.index_hi = rank << self.n_local
        self._plans: dict = {}
        self.bytes_sent = 0
        self.stats = {"exchanges": self.layout.n_exchanges, "gate_steps": sum(1 for s in self.layout.steps if s.kind == "gates")}

    def _device_plan(self, i: int, dtype: torch.dtype):
        from .engine import DevicePlan
        from .plan import TileConfig, build_plan

        key = (i, dtype)
        if key not in self._plans:
            cfg = TileConfig.default(dtype == torch.complex128, False)
            self._plans[key] = DevicePlan(build_plan(self.layout.steps[i].ops, self.n_local, self.slot_of, cfg, n_total=self.n), self.device)
        return self._plans[key]

    def n_sweeps(self, dtype: torch.dtype = torch.complex128) -> int:
        return sum(self._device_plan(i, dtype).n_sweeps for i, s in enumerate(self.layout.steps) if s.kind == "gates")

    def zero_state(self, dtype: torch.dtype = torch.complex128) -> Tensor:
        from . import engine

        st = engine.zero_state(self.n_local, 1, dtype, self.device)
        if self.rank != 0:
            st[0, 0] = 0
        return st

    def run(self, values: Mapping[str, Tensor] | None = None, state: Tensor | None = None, dtype: torch.dtype = torch.complex128) -> Tensor:
        """Returns this rank's `[1, 2^n_local]` shard of U|ψ> in the standard layout (rank = top bits)."""
        from . import cabi, engine

        lib = cabi.load()
        values = values or {}
        st = self.zero_state(dtype) if state is None else state.clone()  # never evolve the caller's tensor in place
        ptable = engine.pack_params(self.names, values, 1, self.device)
        self.bytes_sent = 0
        for i, step in enumerate(self.layout.steps):
            if step.kind == "a2a":
                self.bytes_sent += exchange_global_local(st, self.world, self.group)
                continue
            dp = self._device_plan(i, dtype)
            rc = lib.b200sv_apply_plan(st.data_ptr(), engine._DT[dtype], self.n_local, 1, ptable.data_ptr(), C.byref(dp.struct), 0,
                                       dp.n_sweeps, self.index_hi, dp.workspace(dtype, 1, False).data_ptr(), engine._stream_ptr(self.device))
            cabi.check(rc, "apply_plan (sharded)")
        return st

    def expectation_diagonal(self, state: Tensor, obs: PauliSum) -> Tensor:
        """Re<ψ|O|ψ> for a diagonal (Z-string) observable: local partial sum + all-reduce of one scalar."""
        from . import engine

        if not obs.is_diagonal:
            raise NotImplementedError("only diagonal observables are supported on a qubit-sharded register")
        cps = engine.CompiledPauliSum.build(obs)
        out = self._pauli_expectation(cps, state)
        if self.world > 1:
            dist.all_reduce(out, op=dist.ReduceOp.SUM, group=self.group)
        return out

    def _pauli_expectation(self, cps, state: Tensor) -> Tensor:
        from . import cabi, engine

        dev = state.device
        coef = cps.coef({}, dev)
        # terms are packed for n_total qubits: bit = n-1-q, rank bits resolved through index_hi
        out = torch.zeros(1, dtype=torch.float64, device=dev)
        cf = torch.view_as_real(coef.detach().to(torch.complex128).contiguous())
        rc = cabi.load().b200sv_pauli_expectation(state.data_ptr(), engine._DT[state.dtype], self.n_local, 1, cps.terms(dev).data_ptr(),
                                                  cps.n_terms, cps.n_single, cf.data_ptr(), coef.shape[1], out.data_ptr(), self.index_hi, engine._stream_ptr(dev))
        cabi.check(rc, "pauli_expectation (sharded)")
        return out

    def norm2(self, state: Tensor) -> Tensor:
        from . import engine

        out = engine.dot(state, state).real.clone()
        if self.world > 1:
            dist.all_reduce(out, op=dist.ReduceOp.SUM, group=self.group)
        return out


# ---------------------------------------------------------------------------------------------------
# Global-qubit sharding WITHOUT exchange passes: sweeps address every GPU's shard through NVLink peer memory
# ---------------------------------------------------------------------------------------------------
class _RawDeviceArray:
    """Lets torch view device memory that was not allocated by its caching allocator."""

    def __init__(self, ptr: int, n_float64: int):
        self.__cuda_array_interface__ = {"shape": (n_float64,), "typestr": "<f8", "data": (ptr, False), "version": 2}


class PeerShards:
    """This rank's shard of a `2^n`-amplitude register plus NVLink mappings of every other rank's shard (CUDA IPC).

    `b200sv_ipc_alloc` (cudaMalloc + cudaIpcGetMemHandle) → 64-byte handles all-gathered over the process group →
    `b200sv_ipc_open` on every peer.  `ptrs` is the table `b200sv_apply_plan_peers` takes (own buffer at [rank])."""

    def __init__(self, n_local: int, dtype: torch.dtype, rank: int, world: int, device: torch.device, group=None):
        from . import cabi

        lib = cabi.load()
        self.rank, self.world, self.device, self.dtype, self.group = rank, world, device, dtype, group
        elem = 16 if dtype == torch.complex128 else 8
        self.nbytes = (1 << n_local) * elem
        ptr = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        with torch.cuda.device(device):
            cabi.check(lib.b200sv_ipc_alloc(self.nbytes, C.byref(ptr), handle), "ipc_alloc")
        self.local_ptr = int(ptr.value)
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=device)
        gathered = [torch.empty_like(mine) for _ in range(world)]
        if world > 1:
            dist.all_gather(gathered, mine, group=group)
        else:
            gathered = [mine]
        self._opened: list[int] = []
        ptrs = []
        try:
            for r in range(world):
                if r == rank:
                    ptrs.append(self.local_ptr)
                    continue
                hb = (C.c_ubyte * 64)(*gathered[r].cpu().tolist())
                p = C.c_void_p()
                with torch.cuda.device(device):
                    cabi.check(lib.b200sv_ipc_open(hb, C.byref(p)), "ipc_open")
                self._opened.append(int(p.value))
                ptrs.append(int(p.value))
        except Exception:  # do not leak the mappings opened so far nor the local allocation
            with torch.cuda.device(device):
                for q in self._opened:
                    lib.b200sv_ipc_close(q)
                lib.b200sv_ipc_free(self.local_ptr)
            self.local_ptr, self._opened = None, []
            raise
        self.ptrs = (C.c_void_p * world)(*ptrs)
        real = torch.as_tensor(_RawDeviceArray(self.local_ptr, self.nbytes // 8), device=device)
        pair = real.view(-1, 2) if dtype == torch.complex128 else real.view(torch.float32).view(-1, 2)
        self.state = torch.view_as_complex(pair).reshape(1, -1)  # [1, 2^n_local] view of the local shard

    def close(self) -> None:
        from . import cabi

        lib = cabi.load()
        if self.local_ptr is None:
            return
        torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier(group=self.group)
        with torch.cuda.device(self.device):
            for p in self._opened:
                lib.b200sv_ipc_close(p)
            if self.world > 1:
                dist.barrier(group=self.group)
            lib.b200sv_ipc_free(self.local_ptr)
        self.local_ptr, self.state, self._opened = None, None, []


def peer_traffic(plan, n: int, g: int, elem_bytes: int = 16) -> dict:
    """NVLink bytes one rank moves per direction over all sweeps of a peer-memory plan: a sweep whose tile contains k of
    the g rank bits keeps 2^-k of every tile it processes in its own HBM and reads (then writes back) the rest remotely."""
    shard = (1 << (n - g)) * elem_bytes
    remote = 0.0
    ks = []
    for sw in plan.sweeps:
        k = sum(1 for b in list(sw["tile_bits"])[: int(sw["m"])] if int(b) >= n - g)
        ks.append(k)
        remote += shard * (1.0 - 2.0 ** -k)
    return {"global_bits_per_sweep": ks, "nvlink_bytes_per_rank_per_direction": int(remote), "hbm_bytes_per_rank": 2 * shard * len(ks)}


class PeerShardedSimulator:
    """`QubitShardedSimulator` with the exchanges fused into the sweeps.

    The plan is the ordinary single-device plan of the whole `n`-qubit program; each rank runs its slice of every
    sweep's tiles with `b200sv_apply_plan_peers`, which reads and writes the amplitudes wherever they live (own HBM
    or a peer's over NVLink).  The state never leaves the standard layout (rank = top index bits), there is no
    staging buffer and no all-to-all; the only collective is the barrier that separates two sweeps."""

    def __init__(self, program: Program, rank: int, world: int, device: torch.device | str | None = None, group=None,
                 dtype: torch.dtype = torch.complex128):
        from . import engine
        from .engine import DevicePlan
        from .plan import TileConfig, build_plan

        if world & (world - 1) or world > 16:
            raise ValueError("world size must be a power of two (<= 16)")
        self.program = program
        self.n = program.n_qubits
        self.g = world.bit_length() - 1
        self.n_local = self.n - self.g
        self.rank, self.world, self.group, self.dtype = rank, world, group, dtype
        self.device = engine._dev(device)
        self.names = program.param_names
        self.slot_of = {nm: i for i, nm in enumerate(self.names)}
        fold = 3 if dtype == torch.complex128 else 4
        # 2^11-amplitude tiles, generic kernels (`lean=False`: uncontrolled diagonal gates on global qubits stay free
        # "external diagonals" instead of asking for their qubit as a tile bit, and the NVLink traffic stays what
        # profiles/r01_multi_gpu.md measured).  (2^12 tiles are 15 % faster up to n = 31 but slowed one straddling sweep of the
        # 36-qubit run from ~0.4 s to 1.9 s.)
        cfg = TileConfig(m=11 if dtype == torch.complex128 else 12, r=4, low=2, fold=fold, lean=False)
        self.plan = DevicePlan(build_plan(list(program.ops), self.n, self.slot_of, cfg, packer="greedy"), self.device)
        self._plan_adj: DevicePlan | None = None
        self._lam: PeerShards | None = None
        self.shards = PeerShards(self.n_local, dtype, rank, world, self.device, group)
        self._flag = torch.zeros(1, dtype=torch.float32, device=self.device)
        self.time_sweeps = False  # diagnostics: per-sweep CUDA-event times (incl. the barrier) in `sweep_ms`
        self.sweep_ms: list[float] = []
        self.stats = {"exchanges": 0, "sweeps": self.plan.n_sweeps, **peer_traffic(self.plan.plan, self.n, self.g, 16 if dtype == torch.complex128 else 8)}

    def n_sweeps(self) -> int:
        return self.plan.n_sweeps

    def _barrier(self) -> None:
        """Cross-GPU barrier ON THE STREAM: a 1-element NCCL all-reduce (no host synchronisation)."""
        if self.world > 1:
            dist.all_reduce(self._flag, group=self.group)

    def run(self, values: Mapping[str, Tensor] | None = None, ptable: Tensor | None = None) -> Tensor:
        """Returns the `[1, 2^n_local]` local shard (a view of the IPC buffer, standard layout) of U|0…0>."""
        from . import cabi, engine

        lib = cabi.load()
        st = self.shards.state
        cabi.check(lib.b200sv_init_zero_state(st.data_ptr(), engine._DT[self.dtype], self.n_local, 1, engine._stream_ptr(self.device)), "init")
        if self.rank != 0:
            st[0, 0] = 0
        if ptable is None:
            ptable = engine.pack_params(self.names, values or {}, 1, self.device)
        ptable = ptable.detach().contiguous()
        ws = self.plan.workspace(self.dtype, 1, False).data_ptr()
        self._barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(self.plan.n_sweeps + 1)] if self.time_sweeps else None
        if ev:
            ev[0].record()
        for s in range(self.plan.n_sweeps):
            rc = lib.b200sv_apply_plan_peers(self.shards.ptrs, self.world, self.rank, engine._DT[self.dtype], self.n, ptable.data_ptr(),
                                             C.byref(self.plan.struct), s, s + 1, ws, engine._stream_ptr(self.device))
            cabi.check(rc, "apply_plan_peers")
            self._barrier()
            if ev:
                ev[s + 1].record()
        if ev:
            torch.cuda.synchronize(self.device)
            self.sweep_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(self.plan.n_sweeps)]
        return st

    def norm2(self, state: Tensor) -> Tensor:
        from . import engine

        out = engine.dot(state, state).real.clone()
        if self.world > 1:
            dist.all_reduce(out, op=dist.ReduceOp.SUM, group=self.group)
        return out

    # ---- observables and adjoint gradients on the sharded register -------------------------------
    def _compiled(self, obs: PauliSum):
        from . import engine

        if obs.n_qubits != self.n:
            raise ValueError("observable and register sizes differ")
        return engine.CompiledPauliSum.build(obs)

    def sample(self, values: Mapping[str, Tensor] | None, n_shots: int, uniforms: Tensor | None = None, seed: int = 0,
               run: bool = True) -> Tensor:
        """`n_shots` basis indices (int64, the same tensor on every rank) drawn from |ψ|² of the sharded register
        ("all-gather of the per-rank mass"): the ranks all-gather the probability mass of their shards, every
        shot's uniform picks its rank from the cumulative masses, the owning rank draws the index inside its shard with the
        ordinary inverse-CDF sampler (`b200sv_sample`) on the rescaled uniform, and one all-reduce assembles the indices.
        `uniforms` ([n_shots], in [0, 1)) must be the same on every rank (default: a generator seeded with `seed`)."""
        from . import engine

        if run:
            self.run(values)
        dev = self.device
        st = self.shards.state
        if uniforms is None:
            uniforms = torch.rand(n_shots, dtype=torch.float64, generator=torch.Generator().manual_seed(seed))
        u = uniforms.to(device=dev, dtype=torch.float64).reshape(-1)
        mass = engine.dot(st, st).real.reshape(1).clone()
        masses = [torch.empty_like(mass) for _ in range(self.world)]
        if self.world > 1:
            dist.all_gather(masses, mass, group=self.group)
        else:
            masses = [mass]
        cum = torch.cumsum(torch.cat(masses), 0)
        v = u * cum[-1]
        owner = torch.clamp(torch.searchsorted(cum, v, right=True), max=self.world - 1)
        lo = torch.cat([torch.zeros(1, dtype=torch.float64, device=dev), cum[:-1]])
        mine = owner == self.rank
        out = torch.zeros(n_shots, dtype=torch.int64, device=dev)
        k = int(mine.sum().item())
        if k > 0:
            local_u = torch.clamp((v[mine] - lo[self.rank]) / masses[self.rank], min=0.0, max=1.0 - 2.0**-53)
            idx = engine.sample_indices(st, local_u.reshape(1, -1)).reshape(-1)
            out[mine] = idx + (self.rank << self.n_local)
        if self.world > 1:
            dist.all_reduce(out, op=dist.ReduceOp.SUM, group=self.group)
        return out

    def expectation(self, values: Mapping[str, Tensor] | None, obs: PauliSum, run: bool = True) -> Tensor:
        """Re<ψ|O|ψ> of ANY Pauli-sum observable (constant coefficients): every rank sums over the rows of its shard — partner
        amplitudes of X / Y factors on global qubits are read from the peer shards over NVLink — then one all-reduce of a
        scalar.  `run=False` reuses the state left by the last `run`."""
        from . import cabi, engine

        if run:
            self.run(values)
        cps = self._compiled(obs)
        if cps.is_parametric:
            raise NotImplementedError("parametric observable coefficients are not supported on a sharded register")
        dev = self.device
        coef = cps.coef({}, dev)
        cf = torch.view_as_real(coef.detach().to(torch.complex128).contiguous())
        out = torch.zeros(1, dtype=torch.float64, device=dev)
        rc = cabi.load().b200sv_pauli_expectation_peers(self.shards.ptrs, self.world, self.rank, engine._DT[self.dtype], self.n,
                                                        cps.terms(dev).data_ptr(), cps.n_terms, cps.n_single, cf.data_ptr(), out.data_ptr(),
                                                        engine._stream_ptr(dev))
        cabi.check(rc, "pauli_expectation_peers")
        if self.world > 1:
            dist.all_reduce(out, op=dist.ReduceOp.SUM, group=self.group)
        return out

    def expectation_and_grads(self, values: Mapping[str, Tensor] | None, obs: PauliSum, ptable: Tensor | None = None) -> tuple[Tensor, Tensor]:
        """(E [1], dE/dθ [n_params]) by the adjoint method on the sharded register: λ = Oψ into a second set of peer-mapped
        shards, then the sweeps walked backwards on ψ and λ (`b200sv_adjoint_plan_peers`, one barrier per sweep); every
        rank reduces the gradient moments of its own tiles and ONE all-reduce sums the `[n_params]` vector."""
        from . import cabi, engine
        from .engine import DevicePlan
        from .plan import TileConfig, build_plan

        lib = cabi.load()
        dev = self.device
        if ptable is None:
            ptable = engine.pack_params(self.names, values or {}, 1, dev)
        ptable = ptable.detach().contiguous()
        self.run(ptable=ptable)
        e = self.expectation(None, obs, run=False)
        if self._lam is None:
            self._lam = PeerShards(self.n_local, self.dtype, self.rank, self.world, dev, self.group)
        if self._plan_adj is None:
            fold = 3 if self.dtype == torch.complex128 else 4
            cfg = TileConfig(m=10 if self.dtype == torch.complex128 else 11, r=3, low=2, fold=fold, adjoint=True, lean=False)
            self._plan_adj = DevicePlan(build_plan(list(self.program.ops), self.n, self.slot_of, cfg, packer="greedy"), dev)
        cps = self._compiled(obs)
        coef = cps.coef({}, dev)
        cf = torch.view_as_real(coef.detach().to(torch.complex128).contiguous())
        st = engine._stream_ptr(dev)
        rc = lib.b200sv_pauli_apply_peers(self.shards.ptrs, self.world, self.rank, self._lam.state.data_ptr(), engine._DT[self.dtype], self.n,
                                          cps.terms(dev).data_ptr(), cps.n_terms, cps.n_single, cf.data_ptr(), 1.0, 0.0, st)
        cabi.check(rc, "pauli_apply_peers")
        self._barrier()  # every λ shard is complete before a sweep reads it through the peer table
        grads = torch.zeros(max(1, len(self.names)), 1, dtype=torch.float64, device=dev)
        dp = self._plan_adj
        ws = dp.workspace(self.dtype, 1, True).data_ptr()
        for s in range(dp.n_sweeps - 1, -1, -1):
            rc = lib.b200sv_adjoint_plan_peers(self.shards.ptrs, self._lam.ptrs, self.world, self.rank, engine._DT[self.dtype], self.n,
                                               ptable.data_ptr(), grads.data_ptr(), C.byref(dp.struct), s, s + 1, ws, st)
            cabi.check(rc, "adjoint_plan_peers")
            self._barrier()
        g = grads[:, 0].clone()
        if self.world > 1:
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)
        return e, g

    def close(self) -> None:
        if self._lam is not None:
            self._lam.close()
            self._lam = None
        self.shards.close()


# ---------------------------------------------------------------------------------------------------
# `Configuration(shard="qubits")`: the qadence adapter's entry points for one register sharded over the process group
# ---------------------------------------------------------------------------------------------------
_SIMULATORS: dict = {}


def _simulator(program: Program, dtype: torch.dtype, device: torch.device, group=None) -> "PeerShardedSimulator":
    key = (id(program), dtype, device, id(group))
    if key not in _SIMULATORS:
        if not dist.is_initialized():
            raise RuntimeError('Configuration(shard="qubits") needs an initialised torch.distributed process group (one process per GPU)')
        if any(op.kind >= OpKind.MATK for op in program.ops):
            raise NotImplementedError("a qubit-sharded register runs circuits of simple gates (rotations, controlled gates, SWAP, constant 1-qubit matrices)")
        _SIMULATORS[key] = PeerShardedSimulator(program, dist.get_rank(group), dist.get_world_size(group), device, group, dtype)
    return _SIMULATORS[key]


class _ShardedExpectation(torch.autograd.Function):
    """E(θ) of a Pauli-sum observable on the sharded register; backward = the adjoint sweeps over peer memory."""

    @staticmethod
    def forward(ctx, table: Tensor, sim: "PeerShardedSimulator", obs: PauliSum) -> Tensor:  # type: ignore[override]
        if table.requires_grad:
            e, g = sim.expectation_and_grads(None, obs, ptable=table)
            ctx.save_for_backward(g)
        else:
            sim.run(ptable=table)
            e = sim.expectation(None, obs, run=False)
        return e.reshape(1, 1)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gout: Tensor):  # type: ignore[override]
        (g,) = ctx.saved_tensors
        return (gout.reshape(()) * g).reshape(-1, 1), None, None


def sharded_expectation(program: Program, observables: Sequence[PauliSum], values: Mapping[str, Tensor], dtype: torch.dtype = torch.complex128,
                        device: torch.device | str | None = None, group=None) -> Tensor:
    """`[1, n_obs]` expectation values of `program` on a register sharded by its top qubits over `group`, differentiable
    w.r.t. the entries of `values` (batch size 1: the GPUs are spent on ONE state).  Every rank must make the same call."""
    from . import engine

    dev = engine._dev(device)
    sim = _simulator(program, dtype, dev, group)
    table = engine.pack_params(sim.names, values, 1, dev)
    if table.shape[1] != 1:
        raise ValueError("a qubit-sharded register takes a batch of one parameter set")
    return torch.cat([_ShardedExpectation.apply(table, sim, obs) for obs in observables], dim=1)


def sharded_sample(program: Program, values: Mapping[str, Tensor], n_shots: int, dtype: torch.dtype = torch.complex128,
                   device: torch.device | str | None = None, group=None, seed: int = 0) -> Tensor:
    from . import engine

    sim = _simulator(program, dtype, engine._dev(device), group)
    return sim.sample(values, n_shots, seed=seed)
