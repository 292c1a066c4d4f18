"""HamEvo: ψ ← exp(−i t G) ψ without ever forming a dense 2^n × 2^n generator.

The reference lowers `HamEvo` to `pyq.HamiltonianEvolution`
(`/root/reference/qadence/backends/pyqtorch/convert_ops.py:225-269`), which builds the dense
generator and calls `torch.linalg.matrix_exp` (semantics: `blocks/block_to_tensor.py:409-440`).
Here:

* a diagonal Pauli-sum generator (Z / N strings only — e.g. the Rydberg interaction
  Σ C6/r_ij⁶ N_i N_j that `analog/hamiltonian_terms.py:16-45` produces for `AnalogInteraction`)
  is one elementwise kernel with the diagonal evaluated on the fly (`b200sv_diag_evolve`);
* any other Hermitian generator (Pauli sum with drive terms, dense matrix on a few qubits, generic
  block program) uses a matrix-free Lanczos–Krylov propagator whose matvec is a single
  `b200sv_pauli_apply` / `b200sv_apply_dense` pass, converged to `KRYLOV_TOL` and sub-stepped in
  time when the Krylov space would exceed `KRYLOV_MAX_DIM`.
"""

from __future__ import annotations

from typing import Any, Callable, Mapping, Sequence

import numpy as np
import torch
from torch import Tensor

from . import cabi
from . import engine as E
from .program import Op, PauliSum, Program, parse_td_expr

KRYLOV_TOL = 1e-13
KRYLOV_MAX_DIM = 64
KRYLOV_MEM_BYTES = 48 << 30  # basis storage budget
KRYLOV_RHO_PER_STEP = 24.0  # ||G||·|t| handled per time sub-step (Krylov dimension ≈ ρ + O(ρ^(1/3)) stays < 64)


def _time(op: Op, values: Mapping[str, Tensor], batch: int, device: torch.device) -> Tensor:
    if isinstance(op.param, str):
        t = torch.as_tensor(values[op.param]).detach().reshape(-1)
        if t.is_complex():
            t = t.real
        t = t.to(device=device, dtype=torch.float64)
    else:
        t = torch.tensor([float(np.real(op.param))], dtype=torch.float64, device=device)
    return t.expand(batch).contiguous() if t.numel() == 1 else t.contiguous()


def _matvec(op: Op, compiled: Any, values: Mapping[str, Tensor], n: int, device: torch.device) -> Callable[[Tensor], Tensor]:
    g = op.generator
    if isinstance(g, PauliSum):
        return _pauli_matvec(compiled, compiled.coef(values, device), values, device)
    if isinstance(g, Program):
        def mv(v: Tensor) -> Tensor:
            tab = E.pack_params(compiled.names, values, v.shape[0], device)
            return E._run_segments(compiled, v.clone(), tab, values)
        return mv
    bits = [n - 1 - q for q in op.targets]
    if isinstance(g, str):
        m = torch.as_tensor(values[g]).detach().to(torch.complex128)
        m = m.reshape(-1, m.shape[-2], m.shape[-1])
        if m.shape[0] != 1:
            def mvb(v: Tensor) -> Tensor:  # batch-dependent generator matrices: one batch element at a time
                return torch.cat([E.apply_dense(v[b : b + 1].contiguous(), m[b], bits) for b in range(v.shape[0])])
            return mvb
        mat = m[0]
    else:
        mat = torch.as_tensor(np.asarray(g, dtype=np.complex128))
        mat = mat.reshape(mat.shape[-2], mat.shape[-1])
    return lambda v: E.apply_dense(v, mat, bits)


def _pauli_matvec(compiled: Any, coef: Tensor, values: Mapping[str, Tensor] | None, device: torch.device,
                  coefs_split: "tuple[Tensor, Tensor] | None" = None) -> Callable[..., Tensor]:
    """ψ ↦ Hψ for a Pauli-sum generator (the Lanczos loop passes `out=` and writes straight into its basis).  A generator
    with both diagonal and X / Y terms — every Rydberg Hamiltonian — gets its diagonal tabulated ONCE
    (`b200sv_pauli_diagonal`); each matvec then evaluates only the X / Y terms per amplitude (`b200sv_pauli_apply_diag`)."""
    split = compiled.split_diagonal() if compiled.ps.n_qubits <= 26 else None  # the table is 2^n complex numbers per row
    if split is None:
        return lambda v, out=None: E.pauli_apply(compiled, coef, v, out=out)
    dg, off = split
    if coefs_split is None:
        coefs_split = (dg.coef(values or {}, device), off.coef(values or {}, device))
    table = E.pauli_diagonal(dg, coefs_split[0], device)
    coef_off = coefs_split[1]
    return lambda v, out=None: E.pauli_apply_diag(off, coef_off, table, v, out=out)


def generator_apply(op: Op, compiled: Any, psi: Tensor, values: Mapping[str, Tensor], n: int) -> Tensor:
    """G ψ (used for ∂/∂t in the adjoint pass: ∂U/∂t = −i G U)."""
    return _matvec(op, compiled, values, n, psi.device)(psi)


def _norm_bound(op: Op, compiled: Any, values: Mapping[str, Tensor]) -> float | None:
    """Cheap upper bound of ||G|| (sum of |coefficients| / Frobenius norm), host side; None if unknown."""
    g = op.generator
    if isinstance(g, PauliSum):
        if compiled.is_parametric:
            return float(compiled.coef(values, torch.device("cpu")).abs().sum(0).max())
        return float(np.abs(compiled.consts).sum())
    if isinstance(g, np.ndarray):
        return float(np.linalg.norm(np.asarray(g), ord="fro"))
    if isinstance(g, str):
        return float(torch.linalg.matrix_norm(torch.as_tensor(values[g]).detach().to(torch.complex128).reshape(-1, *values[g].shape[-2:])).max())
    return None


_LAMBDA_CACHE: dict = {}


def _td_coefficients(compiled: Any, tsym: str, t: Tensor, values: Mapping[str, Tensor], B: int, dev: torch.device, diff: str | None = None) -> Tensor:
    """[T, B] complex128 coefficient table of a time-dependent generator at times t [B]; with `diff`, its partial
    derivative w.r.t. that symbol.  The symbolic coefficients are evaluated on the host (T·B numbers per step); the
    matrix-free evolution itself runs on the GPU."""
    import sympy

    def val(nm: str) -> np.ndarray:
        v = torch.as_tensor(values[nm]).detach().cpu().reshape(-1)
        v = (v.real if v.is_complex() else v).to(torch.float64).numpy()
        return v if v.size == B else np.broadcast_to(v, (B,))

    rows = []
    tt = t.detach().cpu().numpy()
    for const, names, term in zip(compiled.consts, compiled.names, compiled.ps.terms):
        c = np.full(B, complex(const))
        for nm in names:
            c = c * val(nm)
        key = (term.expr, diff)
        if key not in _LAMBDA_CACHE:
            e = parse_td_expr(term.expr)[0] if term.expr else sympy.Integer(1)
            if diff is not None:
                e = sympy.diff(e, sympy.Symbol(diff))
            syms = sorted(str(s_) for s_ in e.free_symbols)
            _LAMBDA_CACHE[key] = (syms, sympy.lambdify([sympy.Symbol(s_) for s_ in syms], e, modules="numpy"))
        syms, fn = _LAMBDA_CACHE[key]
        args = [tt if s_ == tsym else val(s_) for s_ in syms]
        c = c * np.broadcast_to(np.asarray(fn(*args), dtype=np.float64), (B,))
        rows.append(c)
    return torch.as_tensor(np.stack(rows)).to(dev)


def _td_step(compiled: Any, coef: Tensor, state: Tensor, dt: Tensor, sign: float, n: int) -> Tensor:
    if compiled.diagonal:
        cf = torch.view_as_real(coef.contiguous())
        rc = cabi.load().b200sv_diag_evolve(state.data_ptr(), E._DT[state.dtype], n, state.shape[0], compiled.terms(state.device).data_ptr(),
                                            compiled.n_terms, compiled.n_single, cf.data_ptr(), coef.shape[1], dt.data_ptr(), state.shape[0],
                                            sign, 0, E._stream_ptr(state.device))
        cabi.check(rc, "diag_evolve")
        return state
    return krylov_expm(lambda v, out=None: E.pauli_apply(compiled, coef, v, out=out), state, sign * dt)


def evolve_time_dependent(op: Op, compiled: Any, psi: Tensor, values: Mapping[str, Tensor], n: int, dagger: bool = False) -> Tensor:
    """dψ/dt = −i H(t) ψ on [0, duration]: `op.steps` exponential-midpoint steps ψ ← exp(−i dt H(t_k + dt/2)) ψ, each one
    a Krylov (or on-the-fly diagonal) exponential.  Second order in dt; unitary at any step count."""
    dev = psi.device
    B = psi.shape[0]
    steps = max(1, op.steps)
    dt = _time(op, values, B, dev) / steps
    out = psi
    for k in (range(steps - 1, -1, -1) if dagger else range(steps)):
        coef = _td_coefficients(compiled, op.tsym, (k + 0.5) * dt, values, B, dev)
        out = _td_step(compiled, coef, out, dt, -1.0 if dagger else 1.0, n)
    return out


def adjoint_time_dependent(op: Op, compiled: Any, psi: Tensor, lam: Tensor, values: Mapping[str, Tensor], n: int,
                           grads: Tensor, slot_of: Mapping[str, int]) -> tuple[Tensor, Tensor]:
    """Backward walk through a time-dependent HamEvo: un-applies the steps on ψ and λ and accumulates
    ∂E/∂duration = 2 Im⟨λ|H(T)|ψ(T)⟩ and ∂E/∂s = ∫ 2 Im⟨λ(t)|∂H/∂s|ψ(t)⟩ dt (trapezoid over each step) for every free
    symbol s of the coefficient expressions.  Both are the continuum derivatives, consistent with the discretised
    forward pass to O(dt²)."""
    from .program import time_dependent_symbols

    dev = psi.device
    B = psi.shape[0]
    steps = max(1, op.steps)
    dur = _time(op, values, B, dev)
    dt = dur / steps
    if isinstance(op.param, str):
        coef_T = _td_coefficients(compiled, op.tsym, dur, values, B, dev)
        grads[slot_of[op.param]] += 2.0 * E.dot(lam, E.pauli_apply(compiled, coef_T, psi)).imag
    symbols = [s_ for s_ in time_dependent_symbols(compiled.ps, op.tsym) if s_ in slot_of]

    def overlaps(tm: Tensor) -> list[Tensor]:
        return [2.0 * E.dot(lam, E.pauli_apply(compiled, _td_coefficients(compiled, op.tsym, tm, values, B, dev, diff=s_), psi)).imag
                for s_ in symbols]

    for k in range(steps - 1, -1, -1):
        tm = (k + 0.5) * dt
        after = overlaps(tm)
        coef = _td_coefficients(compiled, op.tsym, tm, values, B, dev)
        psi = _td_step(compiled, coef, psi, dt, -1.0, n)
        lam = _td_step(compiled, coef, lam, dt, -1.0, n)
        before = overlaps(tm)
        for s_, a, b in zip(symbols, after, before):
            grads[slot_of[s_]] += 0.5 * dt * (a + b)
    return psi, lam


def adjoint_generator_params(op: Op, compiled: Any, psi: Tensor, lam: Tensor, values: Mapping[str, Tensor], n: int, grads: Tensor,
                             slot_of: Mapping[str, int], names: list) -> tuple[Tensor, Tensor]:
    """Backward step through U = exp(−i t H(g)) that also yields ∂E/∂g for coefficients g of the Pauli-sum generator:

        ∂U/∂g = −i t ∫_0^1 e^{−i(1−s) t H} (∂H/∂g) e^{−i s t H} ds   ⇒   ∂E/∂g = 2 t ∫_0^1 Im⟨λ(s)|∂H/∂g|ψ(s)⟩ ds,

    ψ(s), λ(s) being the states evolved back from the end of the gate to the fraction s.  Commuting (all-Z) generators have
    a constant integrand; otherwise Gauss–Legendre nodes (as many as the phase t‖H‖ needs for 1e-12) are visited while
    un-applying the evolution piecewise.  Returns (ψ, λ) before the gate."""
    dev = psi.device
    B = psi.shape[0]
    t = _time(op, values, B, dev)
    dcoefs = [E._dcoef(compiled, nm, values, dev, B) for nm in names]

    def accumulate(weight: float) -> None:
        for nm, dc in zip(names, dcoefs):
            grads[slot_of[nm]] += (2.0 * weight) * t * E.dot(lam, E.pauli_apply(compiled, dc, psi)).imag

    def back(frac: float) -> None:  # un-apply the fraction `frac` of the evolution on both states
        nonlocal psi, lam
        if frac <= 0.0:
            return
        part = Op(op.kind, op.targets, op.controls, param="__t_part__", generator=op.generator, label=op.label)
        vals = {**values, "__t_part__": frac * t}
        psi = evolve(part, compiled, psi, vals, n, dagger=True)
        lam = evolve(part, compiled, lam, vals, n, dagger=True)

    if compiled.diagonal:
        accumulate(1.0)
        back(1.0)
        return psi, lam
    # composite Gauss–Legendre rule: panels of phase <= KRYLOV_RHO_PER_STEP each (the integrand oscillates with the phase
    # t‖H‖, so a single rule capped at a fixed order would under-resolve long evolutions), (phase + 8) nodes per panel
    rho = (_norm_bound(op, compiled, values) or 1.0) * float(t.abs().max())
    panels = max(1, int(np.ceil(rho / KRYLOV_RHO_PER_STEP)))
    k = int(max(8, np.ceil(rho / panels) + 8))
    nodes, weights = np.polynomial.legendre.leggauss(k)
    nodes, weights = 0.5 * (nodes + 1.0), 0.5 * weights  # on [0, 1]
    at = 1.0
    for pnl in range(panels - 1, -1, -1):  # walking back from s = 1 to s = 0
        lo, width = pnl / panels, 1.0 / panels
        for s_, w_ in zip(nodes[::-1], weights[::-1]):
            pos = lo + width * float(s_)
            back(at - pos)
            at = pos
            accumulate(float(w_) * width)
    back(at)
    return psi, lam


def trotter_parts(compiled: Any):
    """(diagonal part, drive part, {qubit: ([X term indices], [Y term indices])}) of a Pauli-sum generator whose non-diagonal
    terms are all single-qubit X / Y terms (the drive of a neutral-atom Hamiltonian), else None."""
    dg = [t for t in compiled.ps.terms if all(p[1] == "Z" for p in t.paulis)]
    off = [t for t in compiled.ps.terms if not all(p[1] == "Z" for p in t.paulis)]
    if any(len(t.paulis) != 1 or t.paulis[0][1] not in ("X", "Y") for t in off):
        return None
    if not hasattr(compiled, "_trotter"):
        cd = E.CompiledPauliSum.build(PauliSum(compiled.ps.n_qubits, dg)) if dg else None
        co = E.CompiledPauliSum.build(PauliSum(compiled.ps.n_qubits, off)) if off else None
        per_q: dict = {}
        if co is not None:
            for i, t in enumerate(co.ps.terms):
                q, axis = t.paulis[0]
                per_q.setdefault(q, ([], []))[0 if axis == "X" else 1].append(i)
        prog = None
        if per_q:
            from .program import ProgramBuilder

            b = ProgramBuilder(compiled.ps.n_qubits)
            for q in sorted(per_q):  # exp(-i θ (cos φ X + sin φ Y)) = RZ(φ) RX(2θ) RZ(-φ): applied right to left
                b.RZ(q, f"__trot_nphi_{q}").RX(q, f"__trot_th_{q}").RZ(q, f"__trot_phi_{q}")
            prog = E.CompiledProgram(b.build())
        compiled._trotter = (cd, co, per_q, prog)
    return compiled._trotter


def evolve_trotter(op: Op, compiled: Any, psi: Tensor, values: Mapping[str, Tensor], n: int, dagger: bool = False) -> Tensor:
    """ψ ← [e^{-i D dt/2} Π_q e^{-i K_q dt} e^{-i D dt/2}]^k ψ, k = op.steps, dt = t / k: second-order Trotter (Strang) splitting
    of H = D + Σ_q K_q into its diagonal part D (all-Z strings: `b200sv_diag_evolve`, evaluated on the fly) and the single-qubit
    drive terms K_q = a_q X_q + b_q Y_q (exact per step: one fused sweep of merged rotations).  Time-symmetric, so the
    dagger is the same sequence with −t.  Error O(t dt² ‖[D, K]‖)."""
    parts = trotter_parts(compiled)
    if parts is None:
        raise NotImplementedError('algo_hevo="trotter" needs a generator whose non-diagonal terms are single-qubit X / Y terms')
    cd, co, per_q, prog = parts
    dev = psi.device
    B = psi.shape[0]
    k = max(1, int(op.steps))
    dt = _time(op, values, B, dev) * ((-1.0 if dagger else 1.0) / k)

    def diag(frac: float) -> None:
        if cd is None:
            return
        coef = cd.coef(values, dev)
        cf = torch.view_as_real(coef.detach().to(torch.complex128).contiguous())
        tt = (frac * dt).contiguous()
        rc = cabi.load().b200sv_diag_evolve(psi.data_ptr(), E._DT[psi.dtype], n, B, cd.terms(dev).data_ptr(), cd.n_terms, cd.n_single, cf.data_ptr(),
                                            coef.shape[1], tt.data_ptr(), B, 1.0, 0, E._stream_ptr(dev))
        cabi.check(rc, "diag_evolve (trotter)")

    table = None
    if prog is not None:
        coef = co.coef(values, dev).real  # [T_off, 1 or B]; Hermitian generator: real coefficients
        vals = {}
        zero = torch.zeros(coef.shape[1], dtype=torch.float64, device=dev)
        for q, (xs, ys) in per_q.items():
            a = coef[xs].sum(0) if xs else zero
            bq = coef[ys].sum(0) if ys else zero
            phi = torch.atan2(bq, a)
            vals[f"__trot_phi_{q}"], vals[f"__trot_nphi_{q}"] = phi, -phi
            vals[f"__trot_th_{q}"] = 2.0 * dt * torch.sqrt(a * a + bq * bq)
        table = E.pack_params(prog.names, vals, B, dev)
    out = psi
    diag(0.5)
    for s_ in range(k):
        if prog is not None:
            out = E._run_segments(prog, out, table, {})
        diag(1.0 if s_ < k - 1 else 0.5)
    return out


_PAULI_NP = {"X": np.array([[0, 1], [1, 0]], dtype=np.complex128), "Y": np.array([[0, -1j], [1j, 0]], dtype=np.complex128),
             "Z": np.array([[1, 0], [0, -1]], dtype=np.complex128)}


def _pauli_terms_dense(ps: PauliSum, support: Sequence[int]) -> np.ndarray:
    """[T, 2^k, 2^k]: the Pauli string of every term on `support` (support[0] = most significant bit), coefficient 1."""
    pos = {q: i for i, q in enumerate(support)}
    out = []
    for term in ps.terms:
        facs = [np.eye(2, dtype=np.complex128)] * len(support)
        for q, p in term.paulis:
            if q not in pos:
                raise ValueError(f"generator term acts on qubit {q} outside the HamEvo support {tuple(support)}")
            facs[pos[q]] = _PAULI_NP[p]
        m = np.ones((1, 1), dtype=np.complex128)
        for f in facs:
            m = np.kron(m, f)
        out.append(m)
    return np.stack(out) if out else np.zeros((0, 2 ** len(support), 2 ** len(support)), dtype=np.complex128)


def lindblad_dissipator(jumps: Sequence[np.ndarray]) -> np.ndarray:
    """Σ_k L_k ⊗ L_k* − ½ (L_k†L_k ⊗ 1 + 1 ⊗ (L_k†L_k)ᵀ) on vec(ρ) (row-major: ρ_ij at index i·d + j, so A ρ B ↦ A ⊗ Bᵀ)."""
    d = np.shape(jumps[0])[0]
    eye = np.eye(d, dtype=np.complex128)
    out = np.zeros((d * d, d * d), dtype=np.complex128)
    for L in jumps:
        L = np.asarray(L, dtype=np.complex128)
        LL = L.conj().T @ L
        out += np.kron(L, L.conj()) - 0.5 * (np.kron(LL, eye) + np.kron(eye, LL.T))
    return out


def evolve_lindblad(op: Op, compiled: Any, rho: Tensor, values: Mapping[str, Tensor], n2: int) -> Tensor:
    """vec(ρ) ← exp(t 𝓛) vec(ρ), 𝓛 = −i (H ⊗ 1 − 1 ⊗ Hᵀ) + dissipator(`op.jumps`): the reference's
    `HamEvo(generator, t, noise_operators=[...])` (backends/pyqtorch/convert_ops.py:249-267 hands the jump operators to
    pyqtorch's master-equation solver).  𝓛 lives on the 2k qubits `op.targets` = (support rows, support columns) of the
    doubled register — 4^k × 4^k, built and exponentiated on the host (k ≤ 5; scipy's Padé `expm`, exact to rounding, instead of
    an ODE solver) per distinct batch element, then applied by `b200sv_apply_dense`.  Time-dependent generators: the product
    of `op.steps` exponential-midpoint steps.  H must be a Pauli sum, a dense matrix or a matrix value on the support."""
    from scipy.linalg import expm

    k = len(op.targets) // 2
    if k > 5:
        raise NotImplementedError(f"Lindblad HamEvo on {k} qubits: the dense superoperator is limited to 5 support qubits")
    sup = list(op.targets[:k])
    d = 2**k
    dev = rho.device
    B = rho.shape[0]
    cpu = torch.device("cpu")
    t = _time(op, values, B, dev).detach().cpu().numpy()  # [B]
    diss = lindblad_dissipator(op.jumps)
    eye = np.eye(d, dtype=np.complex128)

    def liouvillian(H: np.ndarray) -> np.ndarray:
        return -1j * (np.kron(H, eye) - np.kron(eye, H.T)) + diss

    g = op.generator
    if op.tsym:
        terms = _pauli_terms_dense(compiled.ps, sup)
        steps = max(1, op.steps)
        dt = t / steps
        props = [np.eye(d * d, dtype=np.complex128) for _ in range(B)]
        for s_ in range(steps):
            coef = _td_coefficients(compiled, op.tsym, torch.as_tensor((s_ + 0.5) * dt), values, B, cpu).numpy()  # [T, B]
            for b in range(B):
                props[b] = expm(dt[b] * liouvillian(np.tensordot(coef[:, b], terms, axes=1))) @ props[b]
        same = False
    else:
        if isinstance(g, PauliSum):
            coef = compiled.coef(values, cpu).detach().numpy()  # [T, 1 or B]
            terms = _pauli_terms_dense(compiled.ps, sup)
            hs = [np.tensordot(coef[:, b], terms, axes=1) for b in range(coef.shape[1])]
        elif isinstance(g, str):
            m = torch.as_tensor(values[g]).detach().to(torch.complex128).cpu()
            hs = list(m.reshape(-1, m.shape[-2], m.shape[-1]).numpy())
        elif isinstance(g, np.ndarray):
            hs = [np.asarray(g, dtype=np.complex128).reshape(d, d)]
        else:
            raise NotImplementedError("Lindblad HamEvo needs a Pauli-sum or matrix generator")
        same = len(hs) == 1 and bool(np.all(t == t[0]))
        props = [expm(t[b] * liouvillian(hs[b if len(hs) > 1 else 0])) for b in range(1 if same else B)]
    bits = [n2 - 1 - q for q in op.targets]
    if same:
        return E.apply_dense(rho, props[0], bits)
    return torch.cat([E.apply_dense(rho[b : b + 1].contiguous(), props[b], bits) for b in range(B)])


def evolve(op: Op, compiled: Any, psi: Tensor, values: Mapping[str, Tensor], n: int, dagger: bool = False) -> Tensor:
    if op.jumps:
        if dagger:
            raise NotImplementedError("a Lindblad evolution has no inverse")
        return evolve_lindblad(op, compiled, psi, values, n)
    if op.tsym:
        return evolve_time_dependent(op, compiled, psi, values, n, dagger)
    if op.steps > 0 and isinstance(op.generator, PauliSum):
        return evolve_trotter(op, compiled, psi, values, n, dagger)
    dev = psi.device
    B = psi.shape[0]
    t = _time(op, values, B, dev)
    sign = -1.0 if dagger else 1.0
    g = op.generator
    if isinstance(g, PauliSum) and compiled.diagonal:
        coef = compiled.coef(values, dev)
        cf = torch.view_as_real(coef.detach().to(torch.complex128).contiguous())
        rc = cabi.load().b200sv_diag_evolve(
            psi.data_ptr(), E._DT[psi.dtype], n, B, compiled.terms(dev).data_ptr(), compiled.n_terms, compiled.n_single, cf.data_ptr(),
            coef.shape[1], t.data_ptr(), B, sign, 0, E._stream_ptr(dev),
        )
        cabi.check(rc, "diag_evolve")
        return psi
    mv = _matvec(op, compiled, values, n, dev)
    # time sub-stepping chosen up front from a norm bound, so that each step converges within the basis budget
    bound = _norm_bound(op, compiled, values)
    steps = 1
    if bound is not None:
        tmax = float(t.abs().max())
        steps = max(1, int(np.ceil(bound * tmax / KRYLOV_RHO_PER_STEP)))
    out = psi
    for _ in range(steps):
        out = krylov_expm(mv, out, (sign / steps) * t)
    return out


def krylov_expm(matvec: Callable[[Tensor], Tensor], psi: Tensor, t: Tensor, tol: float = KRYLOV_TOL, depth: int = 0) -> Tensor:
    """exp(−i t H) ψ for Hermitian H given only ψ ↦ Hψ; t is `[B]` float64 (per batch element)."""
    B, N = psi.shape
    dev = psi.device
    state_bytes = psi.numel() * psi.element_size()
    m_max = int(max(4, min(KRYLOV_MAX_DIM, N, KRYLOV_MEM_BYTES // max(1, state_bytes) - 2)))
    nrm = torch.sqrt(E.dot(psi, psi).real)
    safe = torch.where(nrm > 0, nrm, torch.ones_like(nrm))
    V = torch.empty(m_max + 1, B, N, dtype=psi.dtype, device=dev)
    V[0].copy_(psi)
    E.scal((1.0 / safe).to(torch.complex128), V[0])
    alphas = torch.zeros(m_max, B, dtype=torch.float64, device=dev)
    betas = torch.zeros(m_max, B, dtype=torch.float64, device=dev)
    acc = torch.zeros(m_max, B, 4, dtype=torch.float64, device=dev)  # per-step accumulators of b200sv_lanczos_step
    lib = cabi.load()
    n_q = int(N).bit_length() - 1
    import inspect

    takes_out = "out" in inspect.signature(matvec).parameters
    y = None
    converged = False
    m = 0
    for j in range(m_max):
        if takes_out:
            matvec(V[j], out=V[j + 1])
        else:
            V[j + 1].copy_(matvec(V[j]))
        # the rest of the step (three-term recurrence, one re-orthogonalisation pass, normalisation, alpha / beta) in three
        # fused passes over the state: no host arithmetic, no synchronisation
        rc = lib.b200sv_lanczos_step(V[j + 1].data_ptr(), V[j].data_ptr(), V[j - 1].data_ptr() if j > 0 else None,
                                     betas[j - 1].data_ptr() if j > 0 else None, acc[j].data_ptr(), alphas[j].data_ptr(), betas[j].data_ptr(),
                                     E._DT[psi.dtype], n_q, B, E._stream_ptr(dev))
        cabi.check(rc, "lanczos_step")
        m = j + 1
        if m >= min(8, m_max) and (m % 8 == 0 or m == m_max or m == N):
            y, err = _small_expm(alphas[:m], betas[:m], t)
            if float(err.max()) < tol or m == N:  # the only host synchronisation: once every 8 steps
                converged = True
                break
    if not converged:
        if depth > 24:
            raise RuntimeError("Krylov HamEvo failed to converge")
        del V
        half = krylov_expm(matvec, psi, 0.5 * t, tol, depth + 1)
        return krylov_expm(matvec, half, 0.5 * t, tol, depth + 1)
    out = torch.empty_like(psi)
    coeff = (y * nrm[None, :].to(torch.complex128)).contiguous()  # [m, B]
    E.lincomb(coeff, V[:m], out)
    return out


def _small_expm(alphas: Tensor, betas: Tensor, t: Tensor) -> tuple[Tensor, Tensor]:
    """y[m, B] = exp(−i t T_m) e_1 for the batched Lanczos tridiagonal and the a-posteriori error
    |β_m y_m|, on the device (`b200sv_tridiag_expm`: scaled Taylor, no eigendecomposition)."""
    m, B = alphas.shape
    dev = alphas.device
    a = alphas.contiguous()
    b = betas.contiguous()
    tt = t.to(torch.float64).contiguous()
    y = torch.empty(m, B, 2, dtype=torch.float64, device=dev)
    err = torch.empty(B, dtype=torch.float64, device=dev)
    rc = cabi.load().b200sv_tridiag_expm(a.data_ptr(), b.data_ptr(), tt.data_ptr(), m, B, y.data_ptr(), err.data_ptr(), E._stream_ptr(dev))
    cabi.check(rc, "tridiag_expm")
    return torch.view_as_complex(y), err
