"""Execution engine: runs `Program`s on a B200 through the C-ABI (`cabi.py` → `libb200sv.so`).

Host-side counterpart of `pyq.QuantumCircuit.run / sample`, `pyq.Observable.expectation` and
`pyqtorch.differentiation.AdjointExpectation` as they are called from
`/root/reference/qadence/backends/pyqtorch/backend.py:158-310` and
`/root/reference/qadence/engines/torch/differentiable_expectation.py:134-165`.
PyTorch is used for device memory, streams and autograd plumbing only; all state-vector arithmetic
happens in the hand-written CUDA kernels.  No CPU fallback: everything here raises without a GPU.
"""

from __future__ import annotations

import ctypes as C
import os
from collections import Counter
from dataclasses import dataclass, field
from typing import Any, Mapping, Sequence

import numpy as np
import torch
from torch import Tensor

from . import cabi
from .embedding import ParamTable
from .plan import SweepPlan, TileConfig, build_plan, pack_pauli_terms
from .program import Op, OpKind, PauliSum, Program

_DT = {torch.complex64: cabi.C64, torch.complex128: cabi.C128}


def _stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _dev(device: torch.device | str | None) -> torch.device:
    cabi.require_device()
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    d = torch.device(device)
    if d.type != "cuda":
        raise cabi.B200svError(f"b200sv executes on CUDA devices only (got {d}); there is no CPU fallback")
    return d if d.index is not None else torch.device("cuda", torch.cuda.current_device())


# ---------------------------------------------------------------------------------------------------
class DevicePlan:
    """A `SweepPlan` uploaded to the device + the `b200sv_plan` struct the C-ABI takes."""

    def __init__(self, plan: SweepPlan, device: torch.device):
        self.plan = plan
        self.device = device
        self._host_sweeps = np.ascontiguousarray(plan.sweeps)

        def up(a: np.ndarray) -> Tensor:
            raw = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
            if raw.size == 0:
                raw = np.zeros(16, dtype=np.uint8)
            return torch.from_numpy(raw.copy()).to(device)

        self._sweeps = up(plan.sweeps)
        self._rounds = up(plan.rounds)
        self._gates = up(plan.gates)
        self._consts = up(plan.consts)
        self._leaders = up(plan.exec_leader)
        self._lean_fwd = up(plan.lean_tab_fwd)
        self._lean_bwd = up(plan.lean_tab_bwd)
        self._exec_mode = up(plan.exec_mode)
        self._host_tma = np.ascontiguousarray(plan.tma) if len(plan.tma) else np.zeros(1, dtype=plan.tma.dtype)
        self.struct = cabi.PlanStruct(
            self._sweeps.data_ptr(),
            self._rounds.data_ptr(),
            self._gates.data_ptr(),
            self._consts.data_ptr(),
            self._host_sweeps.ctypes.data,
            plan.n_sweeps,
            plan.n_exec_total,
            self._leaders.data_ptr(),
            self._lean_fwd.data_ptr(),
            self._lean_bwd.data_ptr(),
            self._host_tma.ctypes.data,
            self._exec_mode.data_ptr(),
        )
        self._workspaces: dict = {}

    def workspace(self, dtype: torch.dtype, batch: int, adjoint: bool) -> Tensor:
        """Device scratch for the resolved-gate table (+ adjoint moments); cached per (dtype, batch, mode)."""
        key = (dtype, batch, adjoint)
        if key not in self._workspaces:
            nbytes = cabi.load().b200sv_plan_workspace_bytes(C.byref(self.struct), _DT[dtype], batch, int(adjoint))
            if nbytes < 0:
                cabi.check(int(nbytes), "plan_workspace_bytes")
            self._workspaces = {key: torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)}
        return self._workspaces[key]

    @property
    def n_sweeps(self) -> int:
        return self.plan.n_sweeps


def _is_unitary_simple(op: Op) -> bool:
    if op.kind == OpKind.MAT1:
        m = np.asarray(op.matrix, dtype=np.complex128)
        return bool(np.allclose(m.conj().T @ m, np.eye(2), atol=1e-12))
    return True


@dataclass
class CompiledPauliSum:
    ps: PauliSum
    terms_np: np.ndarray
    consts: np.ndarray  # [T] complex
    names: list[tuple[str, ...]]
    diagonal: bool
    n_single: int = 0
    _dev_terms: dict = field(default_factory=dict)
    _dev_const_coef: dict = field(default_factory=dict)

    @staticmethod
    def build(ps: PauliSum) -> "CompiledPauliSum":
        ps = ps.simplified()
        # single-Z strings first: the kernels evaluate them once per thread (see include/b200sv.h)
        single = [t for t in ps.terms if len(t.paulis) == 1 and t.paulis[0][1] == "Z"]
        rest = [t for t in ps.terms if not (len(t.paulis) == 1 and t.paulis[0][1] == "Z")]
        ps = PauliSum(ps.n_qubits, single + rest)
        if len(rest) > 1024:
            raise NotImplementedError("Pauli sums with more than 1024 distinct non-trivial terms are not supported yet")
        terms = pack_pauli_terms([t.paulis for t in ps.terms], ps.n_qubits)
        out = CompiledPauliSum(
            ps, terms, np.array([t.const for t in ps.terms], dtype=np.complex128), [t.names for t in ps.terms], ps.is_diagonal
        )
        out.n_single = len(single)
        return out

    @property
    def n_terms(self) -> int:
        return len(self.ps.terms)

    def split_diagonal(self) -> "tuple[CompiledPauliSum, CompiledPauliSum] | None":
        """(diagonal part, X / Y part) as two compiled sums, or None when either is empty (cached).  Used by the Krylov
        matvec: the diagonal part is tabulated once per evolution (`b200sv_pauli_diagonal`)."""
        if not hasattr(self, "_split"):
            dg = [t for t in self.ps.terms if all(p[1] == "Z" for p in t.paulis)]
            off = [t for t in self.ps.terms if not all(p[1] == "Z" for p in t.paulis)]
            self._split = (CompiledPauliSum.build(PauliSum(self.ps.n_qubits, dg)), CompiledPauliSum.build(PauliSum(self.ps.n_qubits, off))) if dg and off else None
        return self._split

    TABLE_MAX_QUBITS = 24  # 2^24 complex weights = 256 MiB; up to n = 22 the table stays in the 126 MB L2
    TABLE_MIN_AMPLITUDES = 1 << 22  # smaller states: the per-amplitude term loop costs microseconds either way

    def table_split(self) -> "tuple[CompiledPauliSum, CompiledPauliSum | None] | None":
        """(diagonal part to tabulate, X / Y part or None) when the diagonal part holds CONSTANT multi-qubit Z strings — the
        terms the Pauli kernels would otherwise evaluate by one parity per term and amplitude (single-Z sums have their own fast
        path) — and the register is small enough for a table; else None.  Cached."""
        if not hasattr(self, "_table_split"):
            self._table_split = None
            n = self.ps.n_qubits
            dg = [t for t in self.ps.terms if all(p[1] == "Z" for p in t.paulis)]
            off = [t for t in self.ps.terms if not all(p[1] == "Z" for p in t.paulis)]
            if n <= self.TABLE_MAX_QUBITS and sum(1 for t in dg if len(t.paulis) > 1) >= 2 and not any(t.names or t.expr for t in dg):
                self._table_split = (CompiledPauliSum.build(PauliSum(n, dg)), CompiledPauliSum.build(PauliSum(n, off)) if off else None)
                self._tables: dict = {}
        return self._table_split

    def diag_table(self, device: torch.device) -> Tensor:
        """[1, 2^n, 2] float64 weights of the tabulated diagonal part (built once per device)."""
        dg = self.table_split()[0]
        if device not in self._tables:
            self._tables[device] = pauli_diagonal(dg, dg.coef({}, device), device)
        return self._tables[device]

    @property
    def is_parametric(self) -> bool:
        return any(len(nm) > 0 for nm in self.names)

    def terms(self, device: torch.device) -> Tensor:
        if device not in self._dev_terms:
            raw = self.terms_np.view(np.uint8).reshape(-1)
            if raw.size == 0:
                raw = np.zeros(32, dtype=np.uint8)
            self._dev_terms[device] = torch.from_numpy(raw.copy()).to(device)
        return self._dev_terms[device]

    def coef(self, values: Mapping[str, Tensor], device: torch.device) -> Tensor:
        """[T, Bc] complex128 coefficient table (Bc = 1 or batch); differentiable w.r.t. values."""
        if not self.is_parametric:
            if device not in self._dev_const_coef:
                self._dev_const_coef[device] = torch.from_numpy(self.consts.reshape(-1, 1).copy()).to(device)
            return self._dev_const_coef[device]
        rows = []
        bc = 1
        for c, nms in zip(self.consts, self.names):
            v: Tensor = torch.full((1,), complex(c), dtype=torch.complex128, device=device)
            for nm in nms:
                p = torch.as_tensor(values[nm]).to(device).reshape(-1)
                v = v * p
            bc = max(bc, v.numel())
            rows.append(v)
        rows = [r.expand(bc) for r in rows]
        return torch.stack(rows) if rows else torch.zeros(0, 1, dtype=torch.complex128, device=device)


def _dcoef(cps: "CompiledPauliSum", name: str, values: Mapping[str, Tensor], device: torch.device, batch: int) -> Tensor:
    """[T, batch] complex128 table of ∂coef_t/∂values[name] (coef_t = const_t · Π values[names_t])."""
    rows = []
    for c, nms in zip(cps.consts, cps.names):
        total = torch.zeros(batch, dtype=torch.complex128, device=device)
        for i, nm in enumerate(nms):
            if nm != name:
                continue
            v = torch.full((batch,), complex(c), dtype=torch.complex128, device=device)
            for j, other in enumerate(nms):
                if j != i:
                    v = v * torch.as_tensor(values[other]).detach().to(device).reshape(-1)
            total = total + v
        rows.append(total)
    return torch.stack(rows) if rows else torch.zeros(0, batch, dtype=torch.complex128, device=device)


@dataclass
class Segment:
    gates: list[Op] | None = None  # run of simple gates
    op: Op | None = None  # one non-fusable op
    compiled: Any = None  # CompiledPauliSum / CompiledProgram(s) for op.generator / op.subs
    _plans: dict = field(default_factory=dict)


class CompiledProgram:
    """A Program split into fused-gate segments and non-fusable ops, with its parameter slots."""

    def __init__(self, program: Program):
        if program.is_noisy:
            raise ValueError("this program carries digital noise: it runs as a density matrix (qadence_b200.density), "
                             "not as a state vector")
        self.program = program
        self.n_qubits = program.n_qubits
        self.names: list[str] = program.param_names
        self.slot_of = {nm: i for i, nm in enumerate(self.names)}
        self.segments: list[Segment] = []
        self._shift_rules: list[int | None] | None = None
        self._absorb: tuple | None = None  # absorb_trailing_permutation: (CompiledProgram without the closing X / CNOT / SWAP gates, those gates)
        self._absorb_obs: dict = {}
        run: list[Op] = []
        for op in program.ops:
            if op.kind < OpKind.MATK:
                run.append(op)
                continue
            if run:
                self.segments.append(Segment(gates=run))
                run = []
            seg = Segment(op=op)
            if op.kind == OpKind.ADD:
                seg.compiled = [CompiledProgram(s) for s in op.subs]
            elif op.kind == OpKind.HAMEVO:
                g = op.generator
                if isinstance(g, PauliSum):
                    seg.compiled = CompiledPauliSum.build(g)
                elif isinstance(g, Program):
                    seg.compiled = CompiledProgram(g)
            self.segments.append(seg)
        if run:
            self.segments.append(Segment(gates=run))
        self.invertible = all(
            all(_is_unitary_simple(o) for o in s.gates) if s.gates is not None
            else (s.op.kind in (OpKind.HAMEVO, OpKind.MATK) and not s.op.jumps)  # a Lindblad evolution is a semigroup: no inverse walk
            for s in self.segments
        )

    def device_plan(self, seg: Segment, dtype: torch.dtype, adjoint: bool, device: torch.device, cfg: TileConfig | None = None) -> DevicePlan:
        key = (dtype, adjoint, device, None if cfg is None else (cfg.m, cfg.r, cfg.low))
        if key not in seg._plans:
            tc = cfg or TileConfig.default(dtype == torch.complex128, adjoint)
            plan = build_plan(seg.gates, self.n_qubits, self.slot_of, tc)
            if (cfg is None and not adjoint and dtype == torch.complex128 and tc.m in (10, 11) and self.n_qubits > 12
                    and "B200SV_TILE_FWD_C128" not in os.environ and int((plan.sweeps["n_gen"] > 0).sum()) > 0):
                # sweeps with generic-phase entries (controlled / diagonal gates, e.g. a QFT) run 17 % faster with
                # 2^12-amplitude tiles; pure rotation + CNOT sweeps do not care (profiles/r01_multi_gpu.md)
                plan = build_plan(seg.gates, self.n_qubits, self.slot_of, TileConfig(m=12, r=4, low=tc.low, fold=tc.fold))
            seg._plans[key] = DevicePlan(plan, device)
        return seg._plans[key]

    def shift_rule(self, slot: int) -> int | None:
        """Number of terms (2 or 4) of the exact parameter-shift rule of this slot, or None when it has none: the slot
        must feed exactly one RX / RY / RZ / PHASE gate (possibly controlled) of the top-level circuit."""
        if self._shift_rules is None:
            uses: dict[str, list] = {}

            def walk(ops: Sequence[Op], nested: bool) -> None:
                for op in ops:
                    if isinstance(op.param, str):
                        uses.setdefault(op.param, []).append((op.kind, bool(op.controls), nested or bool(op.tsym)))
                    for sub in op.subs or []:
                        walk(sub.ops, True)
                    if isinstance(op.generator, Program):
                        walk(op.generator.ops, True)
                    elif isinstance(op.generator, PauliSum):
                        from .program import time_dependent_symbols

                        for nm in [*op.generator.param_names, *time_dependent_symbols(op.generator, op.tsym)]:
                            uses.setdefault(nm, []).append((None, False, True))
                    elif isinstance(op.generator, str):
                        uses.setdefault(op.generator, []).append((None, False, True))

            walk(self.program.ops, False)
            rules: list[int | None] = []
            for nm in self.names:
                u = uses.get(nm, [])
                rule = None
                if len(u) == 1 and not u[0][2]:
                    kind, controlled, _ = u[0]
                    if kind == OpKind.PHASE:
                        rule = 2
                    elif kind in (OpKind.RX, OpKind.RY, OpKind.RZ):
                        rule = 4 if controlled else 2
                rules.append(rule)
            self._shift_rules = rules
        return self._shift_rules[slot]

    def n_sweeps(self, dtype: torch.dtype = torch.complex128, adjoint: bool = False) -> int:
        tc = TileConfig.default(dtype == torch.complex128, adjoint)
        return sum(build_plan(s.gates, self.n_qubits, self.slot_of, tc).n_sweeps for s in self.segments if s.gates is not None)


# ---------------------------------------------------------------------------------------------------
def infer_batch(values: Mapping[str, Any] | None) -> int:
    """Max leading length of the tensor values (`qadence/backends/utils.py:169-184`)."""
    if isinstance(values, ParamTable) and values.clean:
        return values.batch
    b = 1
    for v in (values or {}).values():
        if isinstance(v, Tensor) and v.dim() in (1, 3):  # [B] scalars or [B, k, k] generator matrices; a bare [k, k] matrix is not a batch
            b = max(b, v.shape[0])
    return b


def pack_params(names: Sequence[str], values: Mapping[str, Tensor], batch: int, device: torch.device) -> Tensor:
    """[n_slots, batch] float64 device table of gate parameters, differentiable w.r.t. `values`."""
    if not names:
        return torch.zeros(1, batch, dtype=torch.float64, device=device)
    if isinstance(values, ParamTable) and values.clean and values.table.device == device:
        tab = values.rows(names)  # device-side embedding: the table is already there
        if tab is not None and tab.shape[1] in (1, batch):
            return tab if tab.shape[1] == batch else tab.expand(-1, batch).contiguous()
    rows = []
    for nm in names:
        if nm not in values:
            raise KeyError(f"parameter '{nm}' missing from param_values")
        v = torch.as_tensor(values[nm])
        if v.dim() > 1 and v.numel() != v.shape[0]:
            rows.append(torch.zeros(batch, dtype=torch.float64))  # matrix-valued (generator) entry: not a scalar slot
            continue
        v = v.reshape(-1)
        if v.is_complex():
            v = v.real
        v = v.to(torch.float64)
        if v.numel() == 1:
            v = v.expand(batch)
        elif v.numel() != batch:
            raise ValueError(f"parameter '{nm}' has batch {v.numel()}, expected 1 or {batch}")
        rows.append(v)
    devs = {r.device for r in rows}
    if len(devs) == 1 and next(iter(devs)).type == "cpu":
        return torch.stack(rows).to(device)  # one H2D copy
    return torch.stack([r.to(device) for r in rows])


def zero_state(n: int, batch: int, dtype: torch.dtype, device: torch.device) -> Tensor:
    st = torch.empty(batch, 2**n, dtype=dtype, device=device)
    cabi.check(cabi.load().b200sv_init_zero_state(st.data_ptr(), _DT[dtype], n, batch, _stream_ptr(device)), "init_zero_state")
    return st


def prepare_state(n: int, batch: int, state: Tensor | None, dtype: torch.dtype, device: torch.device) -> Tensor:
    """Fresh, writable `[batch, 2^n]` device state (zero state, or a copy of the user's state)."""
    if state is None:
        return zero_state(n, batch, dtype, device)
    if state.dim() != 2 or state.shape[1] != 2**n:
        raise ValueError(f"state must have shape (batch, 2**{n}); got {tuple(state.shape)}")
    st = state.detach().to(device=device, dtype=dtype)
    if st.shape[0] == 1 and batch > 1:
        st = st.expand(batch, -1)
    elif st.shape[0] != batch:
        raise ValueError(f"state batch {st.shape[0]} incompatible with parameter batch {batch}")
    return st.clone(memory_format=torch.contiguous_format)


# ---------------------------------------------------------------------------------------------------
def pauli_apply(cps: CompiledPauliSum, coef: Tensor, psi: Tensor, out: Tensor | None = None, acc: Tensor | None = None,
                alpha: complex = 1.0, beta: complex = 0.0, index_hi: int = 0) -> Tensor:
    """out = alpha·(Σ_t coef_t P_t ψ) + beta·acc."""
    n = cps.ps.n_qubits
    B = psi.shape[0]
    dev = psi.device
    out = torch.empty_like(psi) if out is None else out
    cf = torch.view_as_real(coef.detach().to(torch.complex128).contiguous())
    rc = cabi.load().b200sv_pauli_apply(
        psi.data_ptr(), out.data_ptr(), 0 if acc is None else acc.data_ptr(), _DT[psi.dtype], n, B,
        cps.terms(dev).data_ptr(), cps.n_terms, cps.n_single, cf.data_ptr(), coef.shape[1],
        float(np.real(alpha)), float(np.imag(alpha)), float(np.real(beta)), float(np.imag(beta)), index_hi, _stream_ptr(dev),
    )
    cabi.check(rc, "pauli_apply")
    return out


def pauli_diagonal(cps: CompiledPauliSum, coef: Tensor, device: torch.device) -> Tensor:
    """[rows, 2^n] complex128 table of a DIAGONAL Pauli sum, rows = coef.shape[1] (1 or the batch)."""
    n = cps.ps.n_qubits
    rows = int(coef.shape[1])
    out = torch.empty(rows, 1 << n, 2, dtype=torch.float64, device=device)
    cf = torch.view_as_real(coef.detach().to(torch.complex128).contiguous())
    rc = cabi.load().b200sv_pauli_diagonal(cps.terms(device).data_ptr(), cps.n_terms, cps.n_single, cf.data_ptr(), rows, n, rows, out.data_ptr(), 0,
                                           _stream_ptr(device))
    cabi.check(rc, "pauli_diagonal")
    return out


def pauli_apply_diag(off: "CompiledPauliSum | None", coef_off: Tensor | None, diag: Tensor, psi: Tensor, out: Tensor | None = None,
                     acc: Tensor | None = None, alpha: complex = 1.0, beta: complex = 0.0, diag_scale: Tensor | None = None) -> Tensor:
    """out = alpha·(s_b·diag ⊙ ψ + Σ_{X/Y terms} coef_t P_t ψ) + beta·acc  (the Krylov matvec with a tabulated diagonal; the
    adjoint pass's λ = Σ_o ḡ·O_o ψ for observables with a tabulated diagonal part: `diag_scale` = s [B] complex128)."""
    n = int(psi.shape[1]).bit_length() - 1
    B = psi.shape[0]
    dev = psi.device
    out = torch.empty_like(psi) if out is None else out
    nt = 0 if off is None else off.n_terms
    cf = torch.view_as_real(coef_off.detach().to(torch.complex128).contiguous()) if nt else None
    ds = torch.view_as_real(diag_scale.detach().to(torch.complex128).contiguous()) if diag_scale is not None else None
    a, b_ = complex(alpha), complex(beta)
    rc = cabi.load().b200sv_pauli_apply_diag(psi.data_ptr(), out.data_ptr(), acc.data_ptr() if acc is not None else 0, _DT[psi.dtype], n, B,
                                             off.terms(dev).data_ptr() if nt else 0, nt, cf.data_ptr() if nt else 0,
                                             coef_off.shape[1] if nt else 1, diag.data_ptr(), diag.shape[0], ds.data_ptr() if ds is not None else 0,
                                             a.real, a.imag, b_.real, b_.imag, _stream_ptr(dev))
    cabi.check(rc, "pauli_apply_diag")
    return out


def pauli_expectation_diag(off: "CompiledPauliSum | None", coef_off: Tensor | None, diag: Tensor, psi: Tensor) -> Tensor:
    """[B] float64: Re⟨ψ|(diag + Σ_{X/Y terms} coef_t P_t)|ψ⟩ with the diagonal part tabulated (`pauli_diagonal`)."""
    n = int(psi.shape[1]).bit_length() - 1
    B = psi.shape[0]
    dev = psi.device
    out = torch.zeros(B, dtype=torch.float64, device=dev)
    nt = 0 if off is None else off.n_terms
    cf = torch.view_as_real(coef_off.detach().to(torch.complex128).contiguous()) if nt else None
    rc = cabi.load().b200sv_pauli_expectation_diag(psi.data_ptr(), _DT[psi.dtype], n, B, off.terms(dev).data_ptr() if nt else 0, nt,
                                                   cf.data_ptr() if nt else 0, coef_off.shape[1] if nt else 1, diag.data_ptr(), diag.shape[0],
                                                   out.data_ptr(), _stream_ptr(dev))
    cabi.check(rc, "pauli_expectation_diag")
    return out


def pauli_expectation(cps: CompiledPauliSum, coef: Tensor, psi: Tensor, index_hi: int = 0) -> Tensor:
    """[B] float64: Re⟨ψ|Σ_t coef_t P_t|ψ⟩."""
    n = cps.ps.n_qubits
    B = psi.shape[0]
    dev = psi.device
    out = torch.zeros(B, dtype=torch.float64, device=dev)
    cf = torch.view_as_real(coef.detach().to(torch.complex128).contiguous())
    rc = cabi.load().b200sv_pauli_expectation(
        psi.data_ptr(), _DT[psi.dtype], n, B, cps.terms(dev).data_ptr(), cps.n_terms, cps.n_single, cf.data_ptr(), coef.shape[1],
        out.data_ptr(), index_hi, _stream_ptr(dev),
    )
    cabi.check(rc, "pauli_expectation")
    return out


def dot(x: Tensor, y: Tensor) -> Tensor:
    """[B] complex128 ⟨x_b|y_b⟩."""
    B = x.shape[0]
    n = int(x.shape[1]).bit_length() - 1
    out = torch.empty(B, 2, dtype=torch.float64, device=x.device)
    cabi.check(cabi.load().b200sv_dot(x.data_ptr(), y.data_ptr(), _DT[x.dtype], n, B, out.data_ptr(), _stream_ptr(x.device)), "dot")
    return torch.view_as_complex(out)


def axpy(a: Tensor, x: Tensor, y: Tensor) -> None:
    """y_b += a[b]·x_b (a: [B] complex128 on device)."""
    B = x.shape[0]
    n = int(x.shape[1]).bit_length() - 1
    ar = torch.view_as_real(a.to(torch.complex128).contiguous())
    cabi.check(cabi.load().b200sv_axpy(ar.data_ptr(), x.data_ptr(), y.data_ptr(), _DT[x.dtype], n, B, _stream_ptr(x.device)), "axpy")


def scal(a: Tensor, x: Tensor) -> None:
    B = x.shape[0]
    n = int(x.shape[1]).bit_length() - 1
    ar = torch.view_as_real(a.to(torch.complex128).contiguous())
    cabi.check(cabi.load().b200sv_scal(ar.data_ptr(), x.data_ptr(), _DT[x.dtype], n, B, _stream_ptr(x.device)), "scal")


def lincomb(c: Tensor, V: Tensor, out: Tensor) -> None:
    """out_b = Σ_k c[k,b]·V[k,b,:]  (c: [K,B] complex128; V: [K,B,2^n])."""
    K, B = c.shape
    n = int(V.shape[2]).bit_length() - 1
    cr = torch.view_as_real(c.to(torch.complex128).contiguous())
    cabi.check(cabi.load().b200sv_lincomb(cr.data_ptr(), V.data_ptr(), K, out.data_ptr(), _DT[V.dtype], n, B, _stream_ptr(V.device)), "lincomb")


def apply_dense(psi: Tensor, matrix: np.ndarray | Tensor, target_bits: Sequence[int]) -> Tensor:
    n = int(psi.shape[1]).bit_length() - 1
    m = torch.as_tensor(matrix).to(torch.complex128).contiguous().to(psi.device)
    mr = torch.view_as_real(m)
    tb = np.asarray(list(target_bits), dtype=np.int32)
    out = torch.empty_like(psi)
    rc = cabi.load().b200sv_apply_dense(psi.data_ptr(), out.data_ptr(), _DT[psi.dtype], n, psi.shape[0], mr.data_ptr(), tb.ctypes.data, len(tb), _stream_ptr(psi.device))
    cabi.check(rc, "apply_dense")
    return out


def bit_reverse(psi: Tensor) -> Tensor:
    n = int(psi.shape[1]).bit_length() - 1
    out = torch.empty_like(psi)
    cabi.check(cabi.load().b200sv_bit_reverse(psi.data_ptr(), out.data_ptr(), _DT[psi.dtype], n, psi.shape[0], _stream_ptr(psi.device)), "bit_reverse")
    return out


# ---------------------------------------------------------------------------------------------------
def _run_segments(cp: CompiledProgram, state: Tensor, ptable: Tensor, values: Mapping[str, Tensor], index_hi: int = 0) -> Tensor:
    from . import hamevo

    lib = cabi.load()
    dev = state.device
    B = state.shape[0]
    for seg in cp.segments:
        if seg.gates is not None:
            dp = cp.device_plan(seg, state.dtype, False, dev)
            rc = lib.b200sv_apply_plan(state.data_ptr(), _DT[state.dtype], cp.n_qubits, B, ptable.data_ptr(), C.byref(dp.struct), 0, dp.n_sweeps,
                                       index_hi, dp.workspace(state.dtype, B, False).data_ptr(), _stream_ptr(dev))
            cabi.check(rc, "apply_plan")
            continue
        op = seg.op
        if op.kind == OpKind.MATK:
            bits = [cp.n_qubits - 1 - q for q in op.targets]
            state = apply_dense(state, op.matrix, bits)
        elif op.kind == OpKind.ADD:
            acc = torch.zeros_like(state)
            for sub in seg.compiled:
                tmp = state.clone()
                sub_table = pack_params(sub.names, values, B, dev)
                tmp = _run_segments(sub, tmp, sub_table, values, index_hi)
                acc += tmp
            state = acc
        elif op.kind == OpKind.HAMEVO:
            state = hamevo.evolve(op, seg.compiled, state, values, cp.n_qubits, dagger=False)
        else:
            raise NotImplementedError(op.kind)
    return state


def run(cp: CompiledProgram, values: Mapping[str, Tensor] | None = None, state: Tensor | None = None,
        dtype: torch.dtype = torch.complex128, device: torch.device | str | None = None) -> Tensor:
    """ψ_out `[B, 2^n]` = U(values)·ψ_in — `Backend.run` (backends/pyqtorch/backend.py:158-179)."""
    dev = _dev(device)
    values = values or {}
    B = infer_batch(values)
    if state is not None:
        B = max(B, state.shape[0])
    ptable = pack_params(cp.names, values, B, dev)
    state_grad = state is not None and state.requires_grad
    if torch.is_grad_enabled() and cp.invertible and (ptable.requires_grad or state_grad):
        # differentiable wavefunction (pyqtorch's `run` is differentiable by autograd; qadence's Overlap / fidelity
        # training relies on it, /root/reference/tests/qadence/test_overlap.py:304-337): the backward pass is the same
        # adjoint walk as for expectations, with torch's cotangent of the state as λ
        return _RunState.apply(cp, values, state, dtype, dev, B, ptable, state if state_grad else torch.zeros(0, device=dev))
    with torch.no_grad():
        st = prepare_state(cp.n_qubits, B, state, dtype, dev)
        return _run_segments(cp, st, ptable.detach(), values)


class _RunState(torch.autograd.Function):
    """ψ_out = U(θ) ψ_in, differentiable (first order) w.r.t. the gate parameters and the input state.

    For a real loss L torch hands `g = ∂L/∂Re ψ + i ∂L/∂Im ψ` to backward, and dL = Re⟨g|dψ⟩.  Hence
    ∂L/∂θ = Re⟨g|∂U/∂θ ψ_in⟩ = ½ · (2 Re⟨λ|∂U/∂θ|ψ⟩ with λ = g) — half of what the adjoint walk reduces — and the cotangent of
    the input state is U†g, the λ the walk ends with."""

    @staticmethod
    def forward(ctx, cp: CompiledProgram, values: Any, state: Tensor | None, dtype: torch.dtype, device: torch.device,
                batch: int, ptable: Tensor, state_in: Tensor) -> Tensor:
        st = prepare_state(cp.n_qubits, batch, state, dtype, device)
        psi = _run_segments(cp, st, ptable, values)
        detached = values.detach() if isinstance(values, ParamTable) else {k: (v.detach() if isinstance(v, Tensor) else v) for k, v in values.items()}
        ctx.cp, ctx.values, ctx.batch = cp, detached, batch
        ctx.grad_names = None if isinstance(values, ParamTable) else {k for k, v in values.items() if isinstance(v, Tensor) and v.requires_grad}
        ctx.state_shape = None if state is None else tuple(state.shape)
        ctx.state_device, ctx.state_dtype = (None, None) if state is None else (state.device, state.dtype)
        ctx.save_for_backward(ptable, psi)
        return psi

    @staticmethod
    @torch.autograd.function.once_differentiable  # no double-backward rule for the wavefunction: create_graph=True raises instead of returning silent zeros
    def backward(ctx, g: Tensor):
        ptable, psi = ctx.saved_tensors
        lam = g.detach().to(psi.dtype).contiguous().clone()
        grads, _, lam_in = _adjoint_walk(ctx.cp, ctx.values, ctx.batch, ptable, psi.detach().clone(), lam, ctx.grad_names)
        g_state = None
        if ctx.needs_input_grad[7] and ctx.state_shape is not None:
            g_state = lam_in if ctx.state_shape[0] == lam_in.shape[0] else lam_in.sum(0, keepdim=True)
            g_state = g_state.to(device=ctx.state_device, dtype=ctx.state_dtype)
        return None, None, None, None, None, None, (0.5 * grads if ctx.needs_input_grad[6] else None), g_state


class CompiledObservable:
    """A Pauli-sum observable (fast path) or a generic linear program (Re⟨ψ|prog ψ⟩)."""

    def __init__(self, obs: PauliSum | Program):
        self.obs = obs
        self.pauli = CompiledPauliSum.build(obs) if isinstance(obs, PauliSum) else None
        self.program = CompiledProgram(obs) if isinstance(obs, Program) else None
        self.n_qubits = obs.n_qubits

    @property
    def names(self) -> list[str]:
        return self.pauli.ps.param_names if self.pauli is not None else self.program.names

    def apply(self, psi: Tensor, values: Mapping[str, Tensor], out: Tensor | None = None) -> Tensor:
        """λ = O ψ (a new tensor, or `out` for Pauli-sum observables)."""
        if self.pauli is not None:
            if self.uses_table(psi):
                dg, off = self.pauli.table_split()
                return pauli_apply_diag(off, off.coef(values, psi.device) if off is not None else None, self.pauli.diag_table(psi.device), psi, out=out)
            return pauli_apply(self.pauli, self.pauli.coef(values, psi.device), psi, out=out)
        tmp = psi.clone()
        tab = pack_params(self.program.names, values, psi.shape[0], psi.device)
        return _run_segments(self.program, tmp, tab, values)

    def uses_table(self, psi: Tensor) -> bool:
        """Constant multi-qubit Z strings on a large state: tabulated diagonal (`CompiledPauliSum.table_split`)."""
        return (self.pauli is not None and psi.numel() >= CompiledPauliSum.TABLE_MIN_AMPLITUDES and os.environ.get("B200SV_DIAG_TABLE", "1") != "0"
                and self.pauli.table_split() is not None)

    def expectation(self, psi: Tensor, values: Mapping[str, Tensor]) -> Tensor:
        if self.pauli is not None:
            if self.uses_table(psi):
                dg, off = self.pauli.table_split()
                return pauli_expectation_diag(off, off.coef(values, psi.device) if off is not None else None, self.pauli.diag_table(psi.device), psi)
            return pauli_expectation(self.pauli, self.pauli.coef(values, psi.device), psi)
        return dot(psi, self.apply(psi, values)).real


def apply_observable(cob: CompiledObservable, psi: Tensor, values: Mapping[str, Tensor]) -> Tensor:
    """O ψ as a new tensor (the density-matrix path traces it: Tr(O ρ), density.py)."""
    return cob.apply(psi, values)


# ---------------------------------------------------------------------------------------------------
class _AdjointExpectation(torch.autograd.Function):
    """E[b,o] = Re⟨ψ_b|O_o|ψ_b⟩ with the adjoint-method backward pass in fused sweeps.

    One backward pass serves every observable: λ = Σ_o ḡ[b,o]·O_o ψ (linearity), where the reference
    runs one `AdjointExpectation.apply` per observable (differentiable_expectation.py:153-165).
    """

    @staticmethod
    def forward(ctx, cp: CompiledProgram, cobs: list[CompiledObservable], values: dict, state: Tensor | None,
                dtype: torch.dtype, device: torch.device, batch: int, ptable: Tensor) -> Tensor:
        st = prepare_state(cp.n_qubits, batch, state, dtype, device)
        psi = _run_segments(cp, st, ptable, values)
        cols = [o.expectation(psi, values) for o in cobs]
        detached = values.detach() if isinstance(values, ParamTable) else {k: (v.detach() if isinstance(v, Tensor) else v) for k, v in values.items()}
        ctx.cp, ctx.cobs, ctx.values, ctx.batch = cp, cobs, detached, batch
        ctx.state, ctx.dtype = state, dtype
        # names whose tensors asked for a gradient (None = unknown, e.g. a device-embedded table): only those HamEvo
        # generator coefficients get the (more expensive) quadrature in the backward pass
        ctx.grad_names = None if isinstance(values, ParamTable) else {k for k, v in values.items() if isinstance(v, Tensor) and v.requires_grad}
        ctx.psi = psi
        ctx.save_for_backward(ptable)
        return torch.stack(cols, dim=1)

    @staticmethod
    def backward(ctx, grad_out: Tensor):
        (ptable,) = ctx.saved_tensors
        if torch.is_grad_enabled() and (ptable.requires_grad or grad_out.requires_grad):
            # create_graph=True: hand out a gradient that can itself be differentiated (DQC-style losses)
            grads = _AdjointGradient.apply(ctx.cp, ctx.cobs, ctx.values, ctx.state, ctx.dtype, ctx.batch, ptable, grad_out)
        else:
            # the adjoint walk un-applies the circuit IN PLACE on the saved wavefunction (no 2^n·B copy per step); a second
            # backward pass over a retained graph finds it consumed and recomputes it
            psi, ctx.psi = ctx.psi, None
            if psi is None:
                with torch.no_grad():
                    st = prepare_state(ctx.cp.n_qubits, ctx.batch, ctx.state, ctx.dtype, ptable.device)
                    psi = _run_segments(ctx.cp, st, ptable.detach(), ctx.values)
            grads = _adjoint_gradient(ctx.cp, ctx.cobs, ctx.values, ctx.batch, ptable, grad_out, psi, ctx.grad_names)
        return None, None, None, None, None, None, None, grads


def _adjoint_gradient(cp: CompiledProgram, cobs: Sequence[CompiledObservable], values: Mapping[str, Any], B: int, ptable: Tensor,
                      grad_out: Tensor, psi: Tensor, grad_names: set | None = None) -> Tensor:
    """`[n_slots, B]` vector-Jacobian product Σ_o ḡ[b,o]·∂E[b,o]/∂θ[slot,b] by the adjoint method: λ = Σ_o ḡ·O_o ψ, then
    one backward walk (`_adjoint_walk`) that un-applies the circuit on ψ and λ and reduces every gate's gradient in the same sweeps.
    CONSUMES `psi` (un-applied in place)."""
    dev = psi.device
    with torch.no_grad():
        if not cp.invertible:
            raise NotImplementedError(
                "adjoint differentiation needs an invertible circuit (unitary gates, Scale, HamEvo); "
                "use diff_mode='gpsr' for circuits with projectors / AddBlocks"
            )
        # λ = Σ_o ḡ[:,o] · O_o ψ
        lam = None
        go = grad_out.to(torch.float64)
        for o, cob in enumerate(cobs):
            w = go[:, o].to(torch.complex128)
            if cob.pauli is not None and cob.uses_table(psi):
                dg, off = cob.pauli.table_split()
                coef = None
                if off is not None:
                    coef = off.coef(values, dev)
                    coef = coef.expand(coef.shape[0], B) * w[None, :]
                new = torch.empty_like(psi)
                pauli_apply_diag(off, coef, cob.pauli.diag_table(dev), psi, out=new, acc=lam, alpha=1.0, beta=0.0 if lam is None else 1.0, diag_scale=w)
                lam = new
            elif cob.pauli is not None:
                coef = cob.pauli.coef(values, dev)
                coef = coef.expand(coef.shape[0], B) * w[None, :]
                if lam is None:
                    lam = pauli_apply(cob.pauli, coef, psi)
                else:
                    new = torch.empty_like(lam)
                    pauli_apply(cob.pauli, coef, psi, out=new, acc=lam, alpha=1.0, beta=1.0)
                    lam = new
            else:
                t = cob.apply(psi, values)
                scal(w, t)
                lam = t if lam is None else lam + t
    return _adjoint_walk(cp, values, B, ptable, psi, lam, grad_names)[0]


def _adjoint_walk(cp: CompiledProgram, values: Mapping[str, Any], B: int, ptable: Tensor, psi: Tensor, lam: Tensor,
                  grad_names: set | None = None) -> tuple[Tensor, Tensor, Tensor]:
    """One backward walk over the circuit: un-applies it on ψ (the circuit's OUTPUT state) and λ (any cotangent state) in
    place / segment by segment and reduces g[slot, b] = 2 Re⟨λ|∂U/∂θ_slot|ψ⟩ for every gate parameter in the same sweeps.
    Returns (g, ψ_in, U†λ).  Linear in λ: λ = Oψ gives ∂⟨O⟩/∂θ, λ = ∂L/∂ψ (torch's cotangent) gives 2·∂L/∂θ."""
    from . import hamevo

    dev = psi.device
    lib = cabi.load()
    with torch.no_grad():
        if not cp.invertible:
            raise NotImplementedError(
                "adjoint differentiation needs an invertible circuit (unitary gates, Scale, HamEvo); "
                "use diff_mode='gpsr' for circuits with projectors / AddBlocks"
            )
        grads = torch.zeros(ptable.shape, dtype=torch.float64, device=dev)
        for seg in reversed(cp.segments):
            if seg.gates is not None:
                dp = cp.device_plan(seg, psi.dtype, True, dev)
                rc = lib.b200sv_adjoint_plan(psi.data_ptr(), lam.data_ptr(), _DT[psi.dtype], cp.n_qubits, B, ptable.data_ptr(),
                                             grads.data_ptr(), C.byref(dp.struct), 0, dp.n_sweeps, 0,
                                             dp.workspace(psi.dtype, B, True).data_ptr(), _stream_ptr(dev))
                cabi.check(rc, "adjoint_plan")
                continue
            op = seg.op
            if op.kind == OpKind.HAMEVO and op.tsym:
                psi, lam = hamevo.adjoint_time_dependent(op, seg.compiled, psi, lam, values, cp.n_qubits, grads, cp.slot_of)
            elif op.kind == OpKind.HAMEVO:
                if isinstance(op.param, str):
                    g_psi = hamevo.generator_apply(op, seg.compiled, psi, values, cp.n_qubits)
                    grads[cp.slot_of[op.param]] += 2.0 * dot(lam, g_psi).imag
                gen_names = [nm for nm in (op.generator.param_names if isinstance(op.generator, PauliSum) else [])
                             if grad_names is None or nm in grad_names]
                if gen_names:  # parametric generator coefficients: ∂E/∂g = 2 t ∫ Im<λ(s)|∂H/∂g|ψ(s)> ds along the evolution
                    psi, lam = hamevo.adjoint_generator_params(op, seg.compiled, psi, lam, values, cp.n_qubits, grads, cp.slot_of, gen_names)
                else:
                    if grad_names is not None and not isinstance(op.generator, PauliSum):
                        bad = [nm for nm in (seg.compiled.names if isinstance(op.generator, Program) else [op.generator] if isinstance(op.generator, str) else [])
                               if nm in grad_names]
                        if bad:
                            raise NotImplementedError(f"adjoint gradients w.r.t. the HamEvo generator parameters {bad} need a Pauli-sum generator; use diff_mode='gpsr'")
                    psi = hamevo.evolve(op, seg.compiled, psi, values, cp.n_qubits, dagger=True)
                    lam = hamevo.evolve(op, seg.compiled, lam, values, cp.n_qubits, dagger=True)
            elif op.kind == OpKind.MATK:
                bits = [cp.n_qubits - 1 - q for q in op.targets]
                m = np.asarray(op.matrix, dtype=np.complex128)
                md = m.conj().T
                if np.allclose(m @ md, np.eye(m.shape[0]), atol=1e-12, rtol=0):
                    psi = apply_dense(psi, md, bits)
                else:  # not unitary (noise superoperators of the density-matrix path): ψ is un-applied with M⁻¹, λ takes M†
                    if np.linalg.cond(m) > 1e8:
                        raise NotImplementedError("adjoint differentiation through a (nearly) singular matrix op — a projector, or a "
                                                  "noise channel at its fully mixing point — is not possible; use diff_mode='gpsr'")
                    psi = apply_dense(psi, np.linalg.inv(m), bits)
                lam = apply_dense(lam, md, bits)
            else:
                raise NotImplementedError(op.kind)
    return grads, psi, lam


# parameter-shift rules used ONLY to differentiate the adjoint gradient once more (second order):
#   two eigenvalues, gap 1 (RX, RY, RZ, PHASE, controlled PHASE):  f' = [f(θ+π/2) − f(θ−π/2)] / 2
#   controlled rotations (eigenvalues 0, ±1/2: gaps 1/2 and 1):      f' = d₊[f(θ+π/2) − f(θ−π/2)] − d₋[f(θ+3π/2) − f(θ−3π/2)]
_SQ2 = 2.0 ** 0.5
SHIFT_RULES = {
    2: [(np.pi / 2, 0.5), (-np.pi / 2, -0.5)],
    4: [(np.pi / 2, (_SQ2 + 1) / (4 * _SQ2)), (-np.pi / 2, -(_SQ2 + 1) / (4 * _SQ2)),
        (3 * np.pi / 2, -(_SQ2 - 1) / (4 * _SQ2)), (-3 * np.pi / 2, (_SQ2 - 1) / (4 * _SQ2))],
}


class _AdjointGradient(torch.autograd.Function):
    """g = Jᵀḡ as a differentiable function of (θ, ḡ): second and higher derivatives of expectation values.

    Its own backward (cotangent w of g) needs J·w for ḡ and Σ_k w_k ∂_k(Jᵀḡ) for θ.  Both come from exact parameter
    shifts of the slots k on which w is non-zero — for a DQC loss these are the feature-map gates only — each shifted
    evaluation being one more forward + adjoint pass on the GPU (and itself differentiable, so third order works too).
    The reference reaches higher orders by torch autograd through pyqtorch's dense ops (`ml_tools/models.py:27-35`,
    `create_graph=True`); its own higher-order adjoint is untested (tests/backends/test_adjoint.py:127-248, skipped)."""

    @staticmethod
    def forward(ctx, cp: CompiledProgram, cobs: list, values: Any, state: Tensor | None, dtype: torch.dtype, batch: int,
                ptable: Tensor, grad_out: Tensor) -> Tensor:
        dev = ptable.device
        st = prepare_state(cp.n_qubits, batch, state, dtype, dev)
        psi = _run_segments(cp, st, ptable, values)
        ctx.cp, ctx.cobs, ctx.values, ctx.state, ctx.dtype, ctx.batch = cp, cobs, values, state, dtype, batch
        ctx.save_for_backward(ptable, grad_out)
        return _adjoint_gradient(cp, cobs, values, batch, ptable, grad_out, psi)

    @staticmethod
    def backward(ctx, w: Tensor):
        ptable, grad_out = ctx.saved_tensors
        cp: CompiledProgram = ctx.cp
        active = torch.nonzero(w.detach().abs().amax(dim=1) > 0).reshape(-1).tolist()
        d_theta = torch.zeros_like(ptable) if ctx.needs_input_grad[6] else None
        d_gout = torch.zeros_like(grad_out) if ctx.needs_input_grad[7] else None
        for k in active:
            rule = cp.shift_rule(k)
            if rule is None:
                raise NotImplementedError(
                    f"second-order differentiation w.r.t. parameter '{cp.names[k]}' is not supported: it must feed exactly one "
                    "rotation / phase gate (use gate-level parameter names, or diff_mode='gpsr')")
            unit = torch.zeros_like(ptable)
            unit[k] = 1.0
            for shift, coef in SHIFT_RULES[rule]:
                theta_s = ptable + shift * unit
                if d_gout is not None:  # (J w)[b,o] += coef · w[k,b] · E[b,o](θ + s e_k)
                    e_s = _AdjointExpectation.apply(cp, ctx.cobs, ctx.values, ctx.state, ctx.dtype, ptable.device, ctx.batch, theta_s)
                    d_gout = d_gout + coef * w[k][:, None] * e_s.to(d_gout.dtype)
                if d_theta is not None:  # Σ_k w_k ∂_k g += coef · w[k,b] · g(θ + s e_k)
                    g_s = _AdjointGradient.apply(cp, ctx.cobs, ctx.values, ctx.state, ctx.dtype, ctx.batch, theta_s, grad_out)
                    d_theta = d_theta + coef * w[k][None, :] * g_s
        return None, None, None, None, None, None, d_theta, d_gout


def absorb_trailing_permutation(cp: CompiledProgram, cobs: Sequence["CompiledObservable"]) -> tuple[CompiledProgram, list["CompiledObservable"]]:
    """(cp', observables') with the same expectation values and gradients: the X / CNOT / SWAP gates that close the circuit are
    conjugated into the Pauli-sum observables (`clifford.py`) when that (1) saves a state round trip — the plan of the circuit
    without them has fewer sweeps — and (2) does not make the observable more expensive: no observable may end up with more
    terms that the Pauli kernels evaluate one parity at a time (anything but single-Z terms and CONSTANT Z strings, which are
    tabulated: `CompiledPauliSum.table_split`).  The closing CNOT ladder of cfg2's HEA turns Z_tot into 20 prefix-parity strings
    Z_0⋯Z_q: absorbed thanks to the table (without it: 1.9 ms saved in the sweeps, 10.6 ms lost in <O> and λ = Oψ,
    profiles/r02_lean_kernel.md §9); the same ladder under an observable with parametric coefficients is not.
    `B200SV_ABSORB_PERM=0` disables it, `=force` skips (2).  Cached on `cp`."""
    from . import clifford

    cobs = list(cobs)
    mode = os.environ.get("B200SV_ABSORB_PERM", "1")
    if mode == "0" or not cobs or any(c.pauli is None for c in cobs):
        return cp, cobs
    if cp._absorb is None:
        cp._absorb = (None, [])
        seg = cp.segments[-1] if cp.segments else None
        if seg is not None and seg.gates is not None:
            body, gates = clifford.split_trailing_permutation(cp.program)
            if gates and len(gates) < len(seg.gates):
                tc = TileConfig.default(True, False)
                full_sweeps = build_plan(seg.gates, cp.n_qubits, cp.slot_of, tc).n_sweeps
                body_cp = CompiledProgram(body)
                bseg = body_cp.segments[-1]
                if body_cp.names == cp.names and bseg.gates is not None and build_plan(bseg.gates, cp.n_qubits, body_cp.slot_of, tc).n_sweeps < full_sweeps:
                    cp._absorb = (body_cp, gates)
    body_cp, gates = cp._absorb
    if body_cp is None:
        return cp, cobs
    def slow_terms(co: "CompiledObservable") -> int:
        if co.pauli.table_split() is not None:  # the whole diagonal part is one table read
            return sum(1 for t in co.obs.terms if not all(p[1] == "Z" for p in t.paulis))
        return sum(1 for t in co.obs.terms if not (len(t.paulis) == 1 and t.paulis[0][1] == "Z"))

    out = []
    for c in cobs:
        hit = cp._absorb_obs.get(id(c))
        if hit is None or hit[0] is not c:
            conj = CompiledObservable(clifford.conjugate_observable(c.obs, gates))
            hit = (c, conj if (mode == "force" or slow_terms(conj) <= slow_terms(c)) else None)
            if len(cp._absorb_obs) > 64:
                cp._absorb_obs.clear()
            cp._absorb_obs[id(c)] = hit
        if hit[1] is None:
            return cp, cobs
        out.append(hit[1])
    return body_cp, out


def expectation(cp: CompiledProgram, observables: Sequence[CompiledObservable], values: Mapping[str, Tensor] | None = None,
                state: Tensor | None = None, dtype: torch.dtype = torch.complex128,
                device: torch.device | str | None = None) -> Tensor:
    """`[B, n_obs]` real expectation values, differentiable (adjoint method) w.r.t. every tensor in `values`
    that feeds a gate parameter or an observable coefficient."""
    dev = _dev(device)
    values = values if isinstance(values, ParamTable) else dict(values or {})
    B = infer_batch(values)
    if state is not None:
        B = max(B, state.shape[0])
    cp, cobs = absorb_trailing_permutation(cp, observables)
    ptable = pack_params(cp.names, values, B, dev)
    real_dt = torch.float64 if dtype == torch.complex128 else torch.float32
    grad_on = torch.is_grad_enabled()
    if not (grad_on and ptable.requires_grad):
        with torch.no_grad():
            st = prepare_state(cp.n_qubits, B, state, dtype, dev)
            psi = _run_segments(cp, st, ptable, values)
            out = torch.stack([o.expectation(psi, values) for o in cobs], dim=1)
            obs_grad = grad_on and any(
                isinstance(values.get(nm), Tensor) and values[nm].requires_grad for o in cobs for nm in o.names
            )
        if not obs_grad:
            return out.to(real_dt)
        return (out + _observable_param_term(cobs, psi, values)).to(real_dt)
    out = _AdjointExpectation.apply(cp, cobs, values, state, dtype, dev, B, ptable)
    if any(isinstance(values.get(nm), Tensor) and values[nm].requires_grad for o in cobs for nm in o.names):
        # recompute ψ is avoided: the Function keeps it; add the (value-free) observable-parameter path
        with torch.no_grad():  # the wavefunction itself is not differentiated here
            psi = run(cp, values, state, dtype, dev)
        out = out + _observable_param_term(cobs, psi, values)
    return out.to(real_dt)


def _observable_param_term(cobs: Sequence[CompiledObservable], psi: Tensor, values: Mapping[str, Tensor]) -> Tensor:
    """Zero-valued tensor carrying dE/d(observable coefficient) = ⟨ψ|P_t|ψ⟩ through autograd."""
    dev = psi.device
    cols = []
    for cob in cobs:
        if cob.pauli is None or not cob.pauli.is_parametric:
            cols.append(torch.zeros(psi.shape[0], dtype=torch.float64, device=dev))
            continue
        cps = cob.pauli
        coef = cps.coef(values, dev)  # differentiable [T, Bc]
        with torch.no_grad():
            per_term = []
            for t in range(cps.n_terms):
                one = torch.zeros(cps.n_terms, 1, dtype=torch.complex128, device=dev)
                one[t, 0] = 1.0
                per_term.append(pauli_expectation(cps, one, psi))
            e = torch.stack(per_term)  # [T, B]
        lin = (coef.real * e).sum(0)
        cols.append(lin - lin.detach())
    return torch.stack(cols, dim=1)


# ---------------------------------------------------------------------------------------------------
def sample_indices(psi: Tensor, uniforms: Tensor) -> Tensor:
    """int64 `[B, n_shots]` basis indices by inverse-CDF on the supplied uniforms (sum-tree descent)."""
    B, N = psi.shape
    n = int(N).bit_length() - 1
    dev = psi.device
    u = uniforms.to(device=dev, dtype=torch.float64).contiguous()
    if u.shape[0] != B:
        raise ValueError("uniforms must have shape [batch, n_shots]")
    shots = u.shape[1]
    chunk = min(N, 256)
    L = N // chunk
    probs = torch.empty(B, N, dtype=torch.float64, device=dev)
    tree = torch.zeros(B, 2 * L, dtype=torch.float64, device=dev)
    out = torch.empty(B, shots, dtype=torch.int64, device=dev)
    rc = cabi.load().b200sv_sample(psi.data_ptr(), _DT[psi.dtype], n, B, probs.data_ptr(), tree.data_ptr(), u.data_ptr(), shots, out.data_ptr(), _stream_ptr(dev))
    cabi.check(rc, "sample")
    return out


def counts_from_indices(idx: Tensor, n: int, little_endian: bool = False) -> list[Counter]:
    out = []
    for row in idx.cpu().numpy():
        vals, cnts = np.unique(row, return_counts=True)
        c = Counter()
        for v, k in zip(vals, cnts):
            s = format(int(v), f"0{n}b")
            c[s[::-1] if little_endian else s] = int(k)
        out.append(c)
    return out


def sample(cp: CompiledProgram, values: Mapping[str, Tensor] | None = None, n_shots: int = 1, state: Tensor | None = None,
           dtype: torch.dtype = torch.complex128, device: torch.device | str | None = None,
           uniforms: Tensor | None = None, little_endian: bool = False) -> list[Counter]:
    """`list[Counter]` of big-endian bitstrings — `Backend.sample` (backends/pyqtorch/backend.py:283-310)."""
    with torch.no_grad():
        psi = run(cp, values, state, dtype, device)
    if uniforms is None:
        uniforms = torch.rand(psi.shape[0], n_shots, dtype=torch.float64, device=psi.device)
    idx = sample_indices(psi, uniforms)
    return counts_from_indices(idx, cp.n_qubits, little_endian)
