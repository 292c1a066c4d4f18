"""CPU oracle — TEST INFRASTRUCTURE ONLY (never imported by `qadence_b200/`).

A restatement, on CPU torch tensors, of the algorithm the reference's hot path runs:
qadence's pyqtorch backend (`/root/reference/qadence/backends/pyqtorch/backend.py:158-310`)
driving third-party `pyqtorch==1.7.7` (`/root/reference/pyproject.toml:57`), whose source is
NOT under `/root/reference` and is not installable here (no network).  The published algorithm of
that dependency is restated from its call sites and from qadence's own dense oracle:

* state layout `[2]*n + [batch]`, one `einsum` of a `[2]*2k + [batch]` operator per gate
  (call sites: `backends/pyqtorch/backend.py:176`, `backends/utils.py:11,112-147,252-258`);
* gate matrices exactly as `qadence/blocks/block_to_tensor.py:27-56` (constants), `:107-126`
  (RX/RY/RZ = cos(θ/2) I − i sin(θ/2) G), `:129-144` (U = RZ(ω)·RY(θ)·RZ(φ)), `:147-157` (PHASE),
  `:175-204` (controlled = P0⊗I + P1⊗U on all-ones controls), `:207-239` (SWAP/CSWAP),
  `:384-407` (Chain = ordered product, Add = sum), `:409-440` (HamEvo = matrix_exp(−i t G)),
  `:455-459` (Scale);
* expectation = Re⟨ψ|O|ψ⟩ with O applied term by term (`backends/pyqtorch/backend.py:181-210`);
* adjoint differentiation: forward ψ = Uψ0, λ = Oψ; backward over gates in reverse,
  ψ ← U†ψ, g_θ = 2 Re⟨λ| ∂U/∂θ |ψ⟩, λ ← U†λ
  (contract at `engines/torch/differentiable_expectation.py:134-165`, behaviour pinned by
  `/root/reference/tests/backends/test_adjoint.py:19-124`);
* sampling: |ψ|² then counting (`backends/pyqtorch/backend.py:283-310`).  The reference draws
  with `torch.multinomial` (RNG stream not pinned by any reference test); the bit-exact contract
  used here is inverse-CDF on SUPPLIED uniforms through a fixed-shape sum tree (see `sample`).

Parity pinning: PINNED.  `tests/test_oracle_golden.py` checks this file against every literal
known-answer vector of `/root/reference/tests/backends/pyq/test_quantum_pyq.py`
(:180-209, :279-316, :394-431, :518-545, :620-632) and `tests/backends/test_endianness.py:113-167`
(committed as `tests/golden/kat_reference.json`), and — whenever `/root/reference` is importable —
against qadence's own dense `block_to_tensor` on every gate family / hea / qft / feature map /
HamEvo (`tests/test_qadence_adapter_cpu.py::test_converted_program_equals_dense_oracle` and the fixture generator
`tests/golden/make_golden.py`, both mirroring `tests/qadence/test_matrices.py:582-771`).

It consumes the same flat IR (`qadence_b200.program`) as the CUDA engine so that golden fixtures
generated here travel to the GPU box.
"""

from __future__ import annotations

import math
import string
from typing import Mapping, Sequence

import numpy as np
import torch

from qadence_b200.program import Op, OpKind, PauliSum, Program, parse_td_expr

CDTYPE = torch.complex128
_LETTERS = string.ascii_letters

_PAULI = {
    "X": torch.tensor([[0, 1], [1, 0]], dtype=CDTYPE),
    "Y": torch.tensor([[0, -1j], [1j, 0]], dtype=CDTYPE),
    "Z": torch.tensor([[1, 0], [0, -1]], dtype=CDTYPE),
}
_I2 = torch.eye(2, dtype=CDTYPE)


# --------------------------------------------------------------------------- layout helpers
def to_pyq_shape(state: torch.Tensor, n: int) -> torch.Tensor:
    """[B, 2^n] → [2]*n + [B]  (qadence/backends/utils.py:112-135)."""
    return state.T.reshape([2] * n + [state.shape[0]])


def from_pyq_shape(state: torch.Tensor) -> torch.Tensor:
    """[2]*n + [B] → [B, 2^n]  (qadence/backends/utils.py:138-147)."""
    return torch.flatten(state, start_dim=0, end_dim=-2).t()


def zero_state(n: int, batch: int) -> torch.Tensor:
    s = torch.zeros(batch, 2**n, dtype=CDTYPE)
    s[:, 0] = 1.0
    return s


def _value(values: Mapping[str, torch.Tensor], p, batch: int) -> torch.Tensor:
    """Resolve a Param to a [batch] tensor (constants broadcast; `[1]` values broadcast)."""
    if isinstance(p, str):
        v = torch.as_tensor(values[p]).reshape(-1)
    else:
        v = torch.as_tensor([p], dtype=CDTYPE if isinstance(p, complex) else torch.float64)
    if v.numel() == 1:
        v = v.expand(batch)
    return v


# --------------------------------------------------------------------------- the einsum kernel
def apply_operator(state: torch.Tensor, operator: torch.Tensor, qubits: Sequence[int]) -> torch.Tensor:
    """ψ ← (O ⊗ I) ψ.  state `[2]*n+[B]`; operator `[2^k, 2^k, B or 1]`, qubits[0] = MSB of the operator."""
    n = state.dim() - 1
    k = len(qubits)
    op = operator.reshape([2] * (2 * k) + [operator.shape[-1]])
    st = list(_LETTERS[: n + 1])  # last letter = batch
    new = list(_LETTERS[n + 1 : n + 1 + k])
    op_sub = "".join(new) + "".join(st[q] for q in qubits) + st[n]
    out = list(st)
    for i, q in enumerate(qubits):
        out[q] = new[i]
    if op.shape[-1] == 1 and state.shape[-1] != 1:
        op = op.expand(*op.shape[:-1], state.shape[-1])
    return torch.einsum(f"{op_sub},{''.join(st)}->{''.join(out)}", op, state)


# --------------------------------------------------------------------------- gate matrices
def rot_matrix(theta: torch.Tensor, axis: str) -> torch.Tensor:
    """cos(θ/2) I − i sin(θ/2) G → [2,2,B]  (block_to_tensor.py:107-126)."""
    theta = theta.to(torch.float64) if not theta.is_complex() else theta
    c = torch.cos(theta / 2).to(CDTYPE)
    s = torch.sin(theta / 2).to(CDTYPE)
    return c[None, None, :] * _I2[:, :, None] - 1j * s[None, None, :] * _PAULI[axis][:, :, None]


def phase_matrix(theta: torch.Tensor) -> torch.Tensor:
    """diag(1, e^{iθ}) → [2,2,B]  (block_to_tensor.py:147-157)."""
    e = torch.exp(1j * theta.to(CDTYPE))
    m = torch.zeros(2, 2, theta.numel(), dtype=CDTYPE)
    m[0, 0] = 1.0
    m[1, 1] = e
    return m


def op_matrix(op: Op, values: Mapping[str, torch.Tensor], batch: int, deriv: bool = False) -> torch.Tensor:
    """The 2x2 target matrix [2,2,B] of a simple 1-qubit op (or its θ-derivative)."""
    if op.kind == OpKind.MAT1:
        assert not deriv
        return torch.as_tensor(op.matrix, dtype=CDTYPE)[:, :, None]
    th = _value(values, op.param, batch)
    if op.kind in (OpKind.RX, OpKind.RY, OpKind.RZ):
        axis = op.kind.name[-1]
        m = rot_matrix(th, axis)
        if deriv:  # d/dθ exp(-iθG/2) = (-i/2) G exp(-iθG/2)
            m = torch.einsum("ij,jkb->ikb", -0.5j * _PAULI[axis], m)
        return m
    if op.kind == OpKind.PHASE:
        m = phase_matrix(th)
        if deriv:
            d = torch.zeros_like(m)
            d[1, 1] = 1j * m[1, 1]
            return d
        return m
    raise ValueError(op.kind)


def controlled(m: torch.Tensor, n_ctrl: int) -> torch.Tensor:
    """P0⊗I + P1⊗U with all-ones controls (block_to_tensor.py:175-204); controls are the MSBs."""
    if n_ctrl == 0:
        return m
    d = 2 ** (n_ctrl + 1)
    out = torch.eye(d, dtype=CDTYPE)[:, :, None].repeat(1, 1, m.shape[-1])
    out[d - 2 :, d - 2 :, :] = m
    return out


def controlled_deriv(m: torch.Tensor, n_ctrl: int) -> torch.Tensor:
    """Derivative of a controlled gate: zero outside the control-active block."""
    if n_ctrl == 0:
        return m
    d = 2 ** (n_ctrl + 1)
    out = torch.zeros(d, d, m.shape[-1], dtype=CDTYPE)
    out[d - 2 :, d - 2 :, :] = m
    return out


_SWAP = torch.tensor([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=CDTYPE)


# --------------------------------------------------------------------------- generators / HamEvo
def paulisum_apply(ps: PauliSum, state: torch.Tensor, values: Mapping[str, torch.Tensor]) -> torch.Tensor:
    """Σ_t coef_t P_t ψ on a pyq-shaped state, one temp state per term (pyq.Add/Scale semantics)."""
    batch = state.shape[-1]
    out = torch.zeros_like(state)
    for t in ps.terms:
        tmp = state
        for q, p in t.paulis:
            tmp = apply_operator(tmp, _PAULI[p][:, :, None], [q])
        coef = torch.full((batch,), t.const, dtype=CDTYPE)
        for nm in t.names:
            coef = coef * _value(values, nm, batch).to(CDTYPE)
        out = out + coef * tmp
    return out


def generator_apply(gen, state: torch.Tensor, values: Mapping[str, torch.Tensor], targets: Sequence[int]) -> torch.Tensor:
    if isinstance(gen, PauliSum):
        return paulisum_apply(gen, state, values)
    if isinstance(gen, Program):
        return _run_pyq(gen, state, values)
    if isinstance(gen, str):
        m = torch.as_tensor(values[gen]).to(CDTYPE)
        m = m.reshape(-1, m.shape[-2], m.shape[-1]).permute(1, 2, 0)  # batch last (convert_ops.py:196-197)
        return apply_operator(state, m, list(targets))
    m = torch.as_tensor(gen, dtype=CDTYPE)
    m = m.reshape(m.shape[-2], m.shape[-1])[:, :, None]
    return apply_operator(state, m, list(targets))


def generator_dense(gen, n: int, values: Mapping[str, torch.Tensor], targets: Sequence[int], batch: int) -> torch.Tensor:
    """Dense [B, 2^n, 2^n] matrix of a generator by applying it to the identity's columns."""
    dim = 2**n
    mats = []
    for b in range(batch):
        vb = {k: (torch.as_tensor(v)[b : b + 1] if torch.as_tensor(v).dim() >= 1 and torch.as_tensor(v).shape[0] == batch and batch > 1 else v) for k, v in values.items()}
        eye = to_pyq_shape(torch.eye(dim, dtype=CDTYPE), n)  # columns as batch
        cols = generator_apply(gen, eye, vb, targets)
        mats.append(from_pyq_shape(cols).T)  # [dim(out), dim(col)]
    return torch.stack(mats)


def td_coefficients(ps: PauliSum, tsym: str, t: torch.Tensor, values, batch: int) -> list[torch.Tensor]:
    """Per-term coefficient [batch] of a time-dependent generator at times t [batch]: const · Π names · expr(t, values)."""
    import sympy

    out = []
    for term in ps.terms:
        c = torch.full((batch,), complex(term.const), dtype=CDTYPE)
        for nm in term.names:
            c = c * _value(values, nm, batch).to(CDTYPE)
        if term.expr:
            e, syms = parse_td_expr(term.expr)
            fn = sympy.lambdify([sympy.Symbol(s_) for s_ in syms], e, modules="numpy")
            args = [t.numpy() if s_ == tsym else _value(values, s_, batch).real.to(torch.float64).numpy() for s_ in syms]
            c = c * torch.as_tensor(np.broadcast_to(np.asarray(fn(*args), dtype=np.float64), (batch,)).copy()).to(CDTYPE)
        out.append(c)
    return out


def hamevo_td_apply(op: Op, state: torch.Tensor, values, n: int, dagger: bool = False) -> torch.Tensor:
    """Time-dependent generator: dψ/dt = −i H(t) ψ on [0, duration] by `op.steps` exponential-midpoint steps
    ψ ← exp(−i dt H(t_k + dt/2)) ψ with exact (dense) exponentials.  The reference hands this case to pyqtorch's ODE
    solvers with `n_steps_hevo` steps (backends/pyqtorch/convert_ops.py:227-231,258-268) and pins it only against
    qutip to 6e-2 (tests/backends/pyq/test_time_dependent_generator.py:25-62)."""
    batch = state.shape[-1]
    dur = _value(values, op.param, batch).to(torch.float64)
    steps = max(1, op.steps)
    dt = dur / steps
    flat = from_pyq_shape(state)
    order = range(steps - 1, -1, -1) if dagger else range(steps)
    for k in order:
        tm = (k + 0.5) * dt
        coefs = td_coefficients(op.generator, op.tsym, tm, values, batch)
        frozen = PauliSum(n, [type(tt)(1.0 + 0j, (), tt.paulis) for tt in op.generator.terms])
        # dense H(t_mid) per batch element
        eye = to_pyq_shape(torch.eye(2**n, dtype=CDTYPE), n)
        H = torch.zeros(batch, 2**n, 2**n, dtype=CDTYPE)
        for term, c in zip(frozen.terms, coefs):
            tmp = eye
            for q, p_ in term.paulis:
                tmp = apply_operator(tmp, _PAULI[p_][:, :, None], [q])
            H = H + c[:, None, None] * from_pyq_shape(tmp).T[None]
        sign = 1j if dagger else -1j
        U = expm_i(H, dt.to(CDTYPE), sign)
        flat = torch.einsum("bij,bj->bi", U, flat)
    return to_pyq_shape(flat, n)


def lindblad_apply(op: Op, state: torch.Tensor, values, n: int) -> torch.Tensor:
    """`HamEvo(generator, t, noise_operators=[L_k])` on a vectorised density matrix (the doubled register of
    qadence_b200/density.py: first n/2 qubits = row index of ρ, last n/2 = column index): the reference hands the jump
    operators to pyqtorch's Lindblad master-equation solver (backends/pyqtorch/convert_ops.py:249-267) and pins the result
    only against qutip `mesolve` to 6e-2 (tests/backends/pyq/test_time_dependent_generator.py:68-126).  Restated with the
    exact semigroup: vec(ρ) ← exp(t 𝓛) vec(ρ) on the FULL register, 𝓛 = −i (H ⊗ 1 − 1 ⊗ Hᵀ) + Σ_k [L_k ⊗ L_k* − ½ (L_k†L_k ⊗ 1 +
    1 ⊗ (L_k†L_k)ᵀ)]; time-dependent generators: `op.steps` exponential-midpoint steps, as `hamevo_td_apply`."""
    from scipy.linalg import expm

    n0 = n // 2
    sup = list(op.targets[: len(op.targets) // 2])
    d = 2**n0
    batch = state.shape[-1]
    eye_state = to_pyq_shape(torch.eye(d, dtype=CDTYPE), n0)
    eye = np.eye(d)
    diss = np.zeros((d * d, d * d), dtype=np.complex128)
    for L in op.jumps:
        Lt = torch.as_tensor(np.asarray(L), dtype=CDTYPE)[:, :, None]
        Lf = from_pyq_shape(apply_operator(eye_state, Lt, sup)).T.numpy()  # the jump operator on the whole register
        LL = Lf.conj().T @ Lf
        diss += np.kron(Lf, Lf.conj()) - 0.5 * (np.kron(LL, eye) + np.kron(eye, LL.T))

    def liouvillian(H: np.ndarray) -> np.ndarray:
        return -1j * (np.kron(H, eye) - np.kron(eye, H.T)) + diss

    t = _value(values, op.param, batch).to(torch.float64).numpy()
    flat = from_pyq_shape(state).numpy().copy()  # [B, d*d]
    if op.tsym:
        steps = max(1, op.steps)
        frozen = [type(tt)(1.0 + 0j, (), tt.paulis) for tt in op.generator.terms]
        dense = []
        for term in frozen:
            tmp = eye_state
            for q, p_ in term.paulis:
                tmp = apply_operator(tmp, _PAULI[p_][:, :, None], [q])
            dense.append(from_pyq_shape(tmp).T.numpy())
        dt = t / steps
        for k in range(steps):
            coefs = td_coefficients(op.generator, op.tsym, torch.as_tensor((k + 0.5) * dt), values, batch)
            for b in range(batch):
                H = sum(complex(c[b]) * m for c, m in zip(coefs, dense))
                flat[b] = expm(dt[b] * liouvillian(H)) @ flat[b]
    else:
        bsz = batch if _generator_is_batched(op, values, batch) else 1
        sub = Op(op.kind, tuple(sup), (), param=op.param, generator=op.generator)
        G = generator_dense(sub.generator, n0, values, sup, bsz).numpy()
        for b in range(batch):
            flat[b] = expm(t[b] * liouvillian(G[b if bsz > 1 else 0])) @ flat[b]
    return to_pyq_shape(torch.as_tensor(flat, dtype=CDTYPE), n)


def expm_i(G: torch.Tensor, t: torch.Tensor, sign: complex) -> torch.Tensor:
    """exp(sign·t·G) for batched G [B, d, d], t [B] (sign = ∓i).

    The reference calls `torch.linalg.matrix_exp` (block_to_tensor.py:409-440), whose Taylor-degree table leaves up to
    2.3e-10 of error on complex128 matrices of small norm (measured: exp(−i x Z), x in (0, 0.3]) — more than the 1e-10
    parity bar, so a correct kernel could fail against it.  For Hermitian G (every generator the reference accepts) the
    oracle therefore exponentiates spectrally, which is exact to rounding; anything else falls back to matrix_exp."""
    me = torch.linalg.matrix_exp(sign * t[:, None, None] * G)
    if torch.allclose(G.detach(), G.detach().conj().transpose(-1, -2), atol=1e-12, rtol=0):
        w, V = torch.linalg.eigh(G.detach())
        phase = torch.exp(sign * t.detach()[:, None].to(CDTYPE) * w.to(CDTYPE))
        exact = (V * phase[:, None, :]) @ V.conj().transpose(-1, -2)
        # value from the spectral form, derivatives from matrix_exp (eigh's backward is singular on degenerate spectra)
        return me + (exact - me.detach())
    return me


def hamevo_apply(op: Op, state: torch.Tensor, values, n: int, dagger: bool = False) -> torch.Tensor:
    """ψ ← exp(−i t G) ψ with a dense exponential (block_to_tensor.py:409-440; see `expm_i` for the one deviation)."""
    if op.jumps:
        if dagger:
            raise ValueError("a Lindblad evolution has no inverse")
        return lindblad_apply(op, state, values, n)
    if op.tsym:
        return hamevo_td_apply(op, state, values, n, dagger)
    if op.steps > 0 and isinstance(op.generator, PauliSum):
        return hamevo_trotter_apply(op, state, values, n, dagger)
    batch = state.shape[-1]
    t = _value(values, op.param, batch).to(CDTYPE)
    gen_depends_on_batch = _generator_is_batched(op, values, batch)
    G = generator_dense(op.generator, n, values, op.targets, batch if gen_depends_on_batch else 1)
    if G.shape[0] == 1:
        G = G.expand(batch, -1, -1)
    sign = 1j if dagger else -1j
    U = expm_i(G, t, sign)
    flat = from_pyq_shape(state)  # [B, dim]
    out = torch.einsum("bij,bj->bi", U, flat)
    return to_pyq_shape(out, n)


def hamevo_trotter_apply(op: Op, state: torch.Tensor, values, n: int, dagger: bool = False) -> torch.Tensor:
    """The b200sv extension `algo_hevo="trotter"` restated with dense exponentials (the reference has no Trotter option:
    its AlgoHEvo values EXP / EIG / RK4, types.py:301-309, are all the exact propagator): `op.steps` Strang steps
    e^{-iD dt/2} e^{-iK dt} e^{-iD dt/2} of H = D + K, D = the all-Z terms, K = the rest."""
    batch = state.shape[-1]
    t = _value(values, op.param, batch).to(CDTYPE)
    g = op.generator
    D = PauliSum(g.n_qubits, [tm for tm in g.terms if all(p[1] == "Z" for p in tm.paulis)])
    K = PauliSum(g.n_qubits, [tm for tm in g.terms if not all(p[1] == "Z" for p in tm.paulis)])
    bsz = batch if _generator_is_batched(op, values, batch) else 1
    GD = generator_dense(D, n, values, op.targets, bsz).expand(batch, -1, -1) if D.terms else None
    GK = generator_dense(K, n, values, op.targets, bsz).expand(batch, -1, -1) if K.terms else None
    k = max(1, int(op.steps))
    sign = 1j if dagger else -1j
    dt = t / k
    half = expm_i(GD, 0.5 * dt, sign) if GD is not None else None
    full = expm_i(GD, dt, sign) if GD is not None else None
    drive = expm_i(GK, dt, sign) if GK is not None else None
    flat = from_pyq_shape(state)

    def mul(U, v):
        return v if U is None else torch.einsum("bij,bj->bi", U, v)

    flat = mul(half, flat)
    for s_ in range(k):
        flat = mul(drive, flat)
        flat = mul(full if s_ < k - 1 else half, flat)
    return to_pyq_shape(flat, n)


def _generator_is_batched(op: Op, values, batch: int) -> bool:
    names: list[str] = []
    g = op.generator
    if isinstance(g, PauliSum):
        names = g.param_names
    elif isinstance(g, Program):
        names = g.param_names
    elif isinstance(g, str):
        v = torch.as_tensor(values[g])
        return v.dim() == 3 and v.shape[0] > 1
    return any(torch.as_tensor(values[nm]).numel() > 1 for nm in names)


# --------------------------------------------------------------------------- program interpreter
def _apply_op(op: Op, state: torch.Tensor, values, n: int, dagger: bool = False) -> torch.Tensor:
    batch = state.shape[-1]
    if op.kind in (OpKind.MAT1, OpKind.RX, OpKind.RY, OpKind.RZ, OpKind.PHASE):
        m = op_matrix(op, values, batch)
        if dagger:
            m = m.conj().transpose(0, 1)
        m = controlled(m, len(op.controls))
        return apply_operator(state, m, [*op.controls, *op.targets])
    if op.kind == OpKind.SWAP:
        m = controlled_k(_SWAP[:, :, None], len(op.controls))
        return apply_operator(state, m, [*op.controls, *op.targets])
    if op.kind == OpKind.SCALE:
        s = _value(values, op.param, batch).to(CDTYPE)
        return state * (s.conj() if dagger else s)
    if op.kind == OpKind.MATK:
        m = torch.as_tensor(op.matrix, dtype=CDTYPE)
        if dagger:
            m = m.conj().T
        return apply_operator(state, m[:, :, None], list(op.targets))
    if op.kind == OpKind.ADD:
        out = torch.zeros_like(state)
        for sub in op.subs:
            out = out + _run_pyq(sub, state, values, dagger=dagger)
        return out
    if op.kind == OpKind.HAMEVO:
        return hamevo_apply(op, state, values, n, dagger=dagger)
    raise ValueError(op.kind)


def controlled_k(m: torch.Tensor, n_ctrl: int) -> torch.Tensor:
    if n_ctrl == 0:
        return m
    k = m.shape[0]
    d = k * 2**n_ctrl
    out = torch.eye(d, dtype=CDTYPE)[:, :, None].repeat(1, 1, m.shape[-1])
    out[d - k :, d - k :, :] = m
    return out


def _run_pyq(program: Program, state: torch.Tensor, values, dagger: bool = False) -> torch.Tensor:
    ops = reversed(program.ops) if dagger else program.ops
    for op in ops:
        state = _apply_op(op, state, values, program.n_qubits, dagger=dagger)
    return state


def infer_batch(values: Mapping[str, torch.Tensor]) -> int:
    """qadence/backends/utils.py:169-184."""
    b = 1
    for v in values.values():
        v = torch.as_tensor(v)
        if v.dim() in (1, 3):
            b = max(b, v.shape[0])
    return b


def run(program: Program, values: Mapping[str, torch.Tensor] | None = None, state: torch.Tensor | None = None) -> torch.Tensor:
    """ψ_out [B, 2^n] = U(values) ψ_in — `Backend.run` (backends/pyqtorch/backend.py:158-179)."""
    values = values or {}
    n = program.n_qubits
    batch = infer_batch(values)
    if state is None:
        st = zero_state(n, batch)
    else:
        st = state.to(CDTYPE)
        if st.shape[0] == 1 and batch > 1:
            st = st.expand(batch, -1)
    out = _run_pyq(program, to_pyq_shape(st, n), values)
    return from_pyq_shape(out).contiguous()


def expectation(program: Program, observables: Sequence[PauliSum | Program], values=None, state=None) -> torch.Tensor:
    """E[b, o] = Re⟨ψ_b|O_o|ψ_b⟩ — `_batched_expectation` (backends/pyqtorch/backend.py:181-210)."""
    values = values or {}
    psi = run(program, values, state)
    n = program.n_qubits
    pyq = to_pyq_shape(psi, n)
    cols = []
    for obs in observables:
        lam = paulisum_apply(obs, pyq, values) if isinstance(obs, PauliSum) else _run_pyq(obs, pyq, values)
        cols.append(torch.einsum("bi,bi->b", psi.conj(), from_pyq_shape(lam)).real)
    return torch.stack(cols, dim=1)


# --------------------------------------------------------------------------- adjoint differentiation
def adjoint_expectation_and_grads(
    program: Program, observable: PauliSum, values: Mapping[str, torch.Tensor], state: torch.Tensor | None = None
) -> tuple[torch.Tensor, dict[str, torch.Tensor]]:
    """Expectation [B] and dE/d(values[name]) as per-batch tensors [B] for every parametric gate.

    Restates the adjoint sweep the reference delegates to
    `pyqtorch.differentiation.AdjointExpectation` (call site
    `engines/torch/differentiable_expectation.py:153-165`): reverse over gates,
    ψ ← U†ψ, g += 2Re⟨λ|∂U ψ⟩, λ ← U†λ.  Gradients for a name used by several gates accumulate.
    Supports simple gates, SCALE (by its real parameter) and HAMEVO time.
    """
    n = program.n_qubits
    psi = to_pyq_shape(run(program, values, state), n)
    batch = psi.shape[-1]
    lam = paulisum_apply(observable, psi, values)
    exp = torch.einsum("...b,...b->b", psi.conj(), lam).real
    grads: dict[str, torch.Tensor] = {}

    def inner(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        return torch.einsum("...b,...b->b", a.conj(), b)

    for op in reversed(program.ops):
        if op.kind == OpKind.SCALE:
            s = _value(values, op.param, batch).to(CDTYPE)
            psi = psi / s
            if isinstance(op.param, str):
                g = 2 * inner(lam, psi).real
                grads[op.param] = grads.get(op.param, 0) + g
            lam = lam * s.conj()
            continue
        psi = _apply_op(op, psi, values, n, dagger=True)
        if isinstance(op.param, str):
            if op.kind == OpKind.HAMEVO:
                mu = -1j * generator_apply(op.generator, _apply_op(op, psi, values, n), values, op.targets)
            else:
                d = controlled_deriv(op_matrix(op, values, batch, deriv=True), len(op.controls))
                mu = apply_operator(psi, d, [*op.controls, *op.targets])
            g = 2 * inner(lam, mu).real
            grads[op.param] = grads.get(op.param, 0) + g
        lam = _apply_op(op, lam, values, n, dagger=True)
    return exp, grads


# --------------------------------------------------------------------------- sampling
SAMPLE_CHUNK = 256  # leaves of the sum tree are sequential sums of 256 consecutive probabilities


def probabilities(state: torch.Tensor) -> np.ndarray:
    """|ψ|² = re·re + im·im with each product and the sum rounded separately (no FMA)."""
    s = state.detach().cpu().numpy()
    re = np.ascontiguousarray(s.real, dtype=np.float64)
    im = np.ascontiguousarray(s.imag, dtype=np.float64)
    return re * re + im * im


def sample_indices(probs: np.ndarray, uniforms: np.ndarray) -> np.ndarray:
    """Inverse-CDF sampling on supplied uniforms through a fixed-shape sum tree.

    probs [B, D] float64 (D = 2^n), uniforms [B, S] in [0,1).  Definition (bit-reproducible on any
    machine because every floating-point addition is fixed):
      leaf_j   = sequential left-to-right sum of probs[j*C : (j+1)*C]        (C = min(D, 256))
      node     = left + right, pairwise up to the root
      v        = u * root; descend from the root: if v < left: go left, else v -= left, go right
      in leaf  : first k with v < running sequential sum (else last index of the leaf)
    Returns int64 indices [B, S] into the 2^n basis (qubit 0 = MSB).
    """
    B, D = probs.shape
    C = min(D, SAMPLE_CHUNK)
    L = D // C
    chunks = probs.reshape(B, L, C)
    run = np.zeros((B, L), dtype=np.float64)
    for k in range(C):  # sequential, fixed order
        run = run + chunks[:, :, k]
    levels = [run]
    while levels[-1].shape[1] > 1:
        a = levels[-1]
        levels.append(a[:, 0::2] + a[:, 1::2])
    out = np.empty(uniforms.shape, dtype=np.int64)
    for b in range(B):
        for s in range(uniforms.shape[1]):
            v = uniforms[b, s] * levels[-1][b, 0]
            node = 0
            for lvl in range(len(levels) - 2, -1, -1):
                left = levels[lvl][b, 2 * node]
                if v < left:
                    node = 2 * node
                else:
                    v = v - left
                    node = 2 * node + 1
            acc = 0.0
            k_sel = C - 1
            base = node * C
            for k in range(C):
                acc = acc + probs[b, base + k]
                if v < acc:
                    k_sel = k
                    break
            out[b, s] = base + k_sel
    return out


def counts_from_indices(idx: np.ndarray, n: int) -> list[dict[str, int]]:
    """Counter-style dicts keyed by big-endian bitstrings (backends/pyqtorch/backend.py:300-302)."""
    out = []
    for row in idx:
        vals, cnts = np.unique(row, return_counts=True)
        out.append({format(int(v), f"0{n}b"): int(c) for v, c in zip(vals, cnts)})
    return out


def invert_endianness_state(state: torch.Tensor, n: int) -> torch.Tensor:
    """Bit-reversal permutation of the basis axis (qadence/transpile/invert.py:95-109)."""
    idx = torch.tensor([int(format(i, f"0{n}b")[::-1], 2) for i in range(2**n)])
    return state[:, idx]
