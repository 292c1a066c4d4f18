"""Density matrices and digital / readout noise on top of the state-vector kernels.

The reference gets mixed states from pyqtorch: a gate that carries `noise=` turns the state into a
`pyqtorch.utils.DensityMatrix`, applies `U ρ U†` and then every channel `ρ → Σ_k K_k ρ K_k†` of its protocol
(`/root/reference/qadence/backends/pyqtorch/convert_ops.py:206-208,354-372`, `backend.py:42-79`); `run` then returns
`[batch, 2^n, 2^n]` (`backends/utils.py:140-147`), `expectation` is `Tr(O ρ)` and `sample` draws from `diag(ρ)`, optionally
corrupted by a readout confusion matrix (`convert_ops.py:375-393`).

Here ρ is **vectorised**: `v[i·2^n + j] = ρ_ij` is an ordinary `[batch, 2^(2n)]` state of a 2n-qubit register whose
qubits 0…n−1 are the row index and n…2n−1 the column index.  Then

* `U ρ U†`            →  `U` on the row qubits and `conj(U)` on the column qubits — two ordinary gates, fused by the
  sweep planner like any others (RX(θ)* = RX(−θ), RY* = RY, RZ(θ)* = RZ(−θ), PHASE(θ)* = PHASE(−θ),
  `exp(−itG)* = exp(−it(−G*))`, constants conjugated);
* a channel on qubit q →  the 4×4 superoperator `Σ_k K_k ⊗ conj(K_k)` on the qubit pair (q, q+n): one constant dense
  op (`b200sv_apply_dense`); successive channels of one gate are pre-multiplied;
* `|ψ⟩⟨ψ|`            →  `ψ ⊗ conj(ψ)`; the zero state is the zero state of the doubled register;
* `Tr(O ρ)`           →  `O` applied to the row qubits (`b200sv_pauli_apply`), then the sum of the diagonal.

So the whole noisy path is a *program transformation* (`vectorise`) in front of the unchanged engine — no new kernel —
and is exercised on CPU by running the transformed program through the oracle.

**Parity status: unpinned.**  The Kraus operators and "noise acts on the gate's first target qubit, after the gate" are
restated from pyqtorch 1.7.7's published behaviour (`pyqtorch/noise/digital_gates.py`, `Primitive` forward); pyqtorch
is not in this image and the reference's own noise tests only check self-consistency (noisy ≠ noiseless, model ==
backend; `tests/qadence/test_noise/test_digital_noise.py:74-127`).  The channels are the textbook ones (Nielsen &
Chuang §8.3) and are checked against an independent dense Kraus-sum evolution in `tests/test_density.py`.
"""

from __future__ import annotations

from collections import Counter
from logging import getLogger
from typing import Any, Mapping, Sequence

import numpy as np
import torch
from torch import Tensor

from . import engine
from .program import CONST_MATS, Op, OpKind, PauliSum, PauliTerm, Program

try:  # the class qadence checks with isinstance (backends/utils.py:119,145; states.py:343)
    from pyqtorch.utils import DensityMatrix  # type: ignore
except Exception:  # pyqtorch absent: same definition (a bare Tensor subclass)

    class DensityMatrix(Tensor):  # type: ignore[no-redef]
        pass


logger = getLogger(__name__)

NEG = "__neg__"  # key prefix of the negated copy of a gate parameter (conjugated rotations)
NEGCONJ = "__negconj__"  # ... of −conj(matrix) of a matrix-valued HamEvo generator entry

_I2, _X, _Y, _Z = CONST_MATS["I"], CONST_MATS["X"], CONST_MATS["Y"], CONST_MATS["Z"]


# ---------------------------------------------------------------------------------------------------
# channels
# ---------------------------------------------------------------------------------------------------
def _probs(p: Any, k: int, name: str) -> tuple[float, ...]:
    vals = tuple(float(x) for x in (p if isinstance(p, (tuple, list, np.ndarray)) else (torch.as_tensor(p).reshape(-1).tolist())))
    if len(vals) != k:
        raise ValueError(f"{name} takes {k} error probabilit{'y' if k == 1 else 'ies'}, got {len(vals)}")
    if any(x < 0 or x > 1 for x in vals):
        raise ValueError(f"{name}: error probabilities must lie in [0, 1], got {vals}")
    return vals


def kraus_operators(protocol: str, error_probability: Any) -> list[np.ndarray]:
    """Kraus operators of a `NoiseProtocol.DIGITAL` member (values of pyqtorch's `DigitalNoiseType`)."""
    name = str(getattr(protocol, "value", protocol))
    if name == "BitFlip":
        (p,) = _probs(error_probability, 1, name)
        return [np.sqrt(1 - p) * _I2, np.sqrt(p) * _X]
    if name == "PhaseFlip":
        (p,) = _probs(error_probability, 1, name)
        return [np.sqrt(1 - p) * _I2, np.sqrt(p) * _Z]
    if name == "Depolarizing":
        (p,) = _probs(error_probability, 1, name)
        return [np.sqrt(1 - p) * _I2, np.sqrt(p / 3) * _X, np.sqrt(p / 3) * _Y, np.sqrt(p / 3) * _Z]
    if name == "PauliChannel":
        px, py, pz = _probs(error_probability, 3, name)
        if px + py + pz > 1:
            raise ValueError("PauliChannel: the probabilities must sum to at most 1")
        return [np.sqrt(1 - px - py - pz) * _I2, np.sqrt(px) * _X, np.sqrt(py) * _Y, np.sqrt(pz) * _Z]
    if name == "AmplitudeDamping":
        (g,) = _probs(error_probability, 1, name)
        return [np.array([[1, 0], [0, np.sqrt(1 - g)]], dtype=np.complex128), np.array([[0, np.sqrt(g)], [0, 0]], dtype=np.complex128)]
    if name == "PhaseDamping":
        (g,) = _probs(error_probability, 1, name)
        return [np.array([[1, 0], [0, np.sqrt(1 - g)]], dtype=np.complex128), np.array([[0, 0], [0, np.sqrt(g)]], dtype=np.complex128)]
    if name == "GeneralizedAmplitudeDamping":
        p, g = _probs(error_probability, 2, name)
        return [
            np.sqrt(p) * np.array([[1, 0], [0, np.sqrt(1 - g)]], dtype=np.complex128),
            np.sqrt(p) * np.array([[0, np.sqrt(g)], [0, 0]], dtype=np.complex128),
            np.sqrt(1 - p) * np.array([[np.sqrt(1 - g), 0], [0, 1]], dtype=np.complex128),
            np.sqrt(1 - p) * np.array([[0, 0], [np.sqrt(g), 0]], dtype=np.complex128),
        ]
    raise NotImplementedError(f"unknown digital noise protocol '{name}'")


N_PROBS = {"BitFlip": 1, "PhaseFlip": 1, "Depolarizing": 1, "PauliChannel": 3, "AmplitudeDamping": 1, "PhaseDamping": 1,
           "GeneralizedAmplitudeDamping": 2}


def channel_superoperator(kraus: Sequence[np.ndarray]) -> np.ndarray:
    """4×4 matrix of ρ → Σ K ρ K† on vec(ρ), basis |row bit, column bit⟩ (row bit most significant)."""
    return sum(np.kron(k, k.conj()) for k in kraus)


def noise_superoperator(noise: Sequence[tuple[str, Sequence[float]]]) -> np.ndarray:
    s = np.eye(4, dtype=np.complex128)
    for name, probs in noise:
        s = channel_superoperator(kraus_operators(name, tuple(probs))) @ s
    return s


# ---------------------------------------------------------------------------------------------------
# program transformation
# ---------------------------------------------------------------------------------------------------
def _lift_paulisum(ps: PauliSum, n2: int, shift: int, negconj: bool) -> PauliSum:
    terms = []
    for t in ps.terms:
        c = complex(t.const)
        if negconj:  # −conj(c · P): conj(Y) = −Y
            c = -np.conj(c) * (-1) ** sum(1 for _, p in t.paulis if p == "Y")
        terms.append(PauliTerm(c, t.names, tuple((q + shift, p) for q, p in t.paulis), t.expr))
    return PauliSum(n2, terms)


def _lift_op(op: Op, n2: int, shift: int, conj: bool, needs: set[str]) -> list[Op]:
    """`op` re-declared on the doubled register: as is on the row qubits (shift 0) or conjugated on the column qubits."""
    targets = tuple(q + shift for q in op.targets)
    controls = tuple(q + shift for q in op.controls)
    k = op.kind
    param = op.param
    if conj and param is not None and k in (OpKind.RX, OpKind.RZ, OpKind.PHASE):
        if isinstance(param, str):
            needs.add(param)
            param = NEG + param
        else:
            param = -param
    if conj and k == OpKind.SCALE and param is not None and not isinstance(param, str):
        param = complex(param).conjugate()
        param = param.real if param.imag == 0 else param
    matrix = None if op.matrix is None else (np.conj(op.matrix) if conj else np.array(op.matrix))
    subs = None if op.subs is None else [lift_program(s, n2, shift, conj, needs) for s in op.subs]
    gen = op.generator
    if k == OpKind.HAMEVO:
        if isinstance(gen, PauliSum):
            gen = _lift_paulisum(gen, n2, shift, conj)
        elif isinstance(gen, Program):
            gen = lift_program(gen, n2, shift, conj, needs)
            if conj:
                gen = Program(n2, [*gen.ops, Op(OpKind.SCALE, param=-1.0, label="Scale")])
        elif isinstance(gen, str):
            if conj:
                needs.add(NEGCONJ + gen)
                gen = NEGCONJ + gen
        elif gen is not None:
            gen = -np.conj(gen) if conj else np.array(gen)
    return [Op(k, targets, controls, param=param, matrix=matrix, subs=subs, generator=gen, label=op.label, tsym=op.tsym, steps=op.steps)]


def lift_program(prog: Program, n2: int, shift: int, conj: bool, needs: set[str]) -> Program:
    out: list[Op] = []
    for op in prog.ops:
        if op.noise:
            raise NotImplementedError("digital noise inside an AddBlock / generator is not supported")
        out.extend(_lift_op(op, n2, shift, conj, needs))
    return Program(n2, out)


def vectorise(prog: Program) -> tuple[Program, list[str]]:
    """Program on n qubits (with per-op `noise`) → equivalent program on vec(ρ) (2n qubits, noise-free ops only), plus the
    parameter names whose negated copies (`NEG + name`, `NEGCONJ + name`) must be added to `param_values`."""
    n = prog.n_qubits
    n2 = 2 * n
    needs: set[str] = set()
    out: list[Op] = []
    for op in prog.ops:
        if op.kind == OpKind.HAMEVO and op.jumps:
            # Lindblad evolution: ONE op on the row + column copies of the support; the generator stays the n-qubit one and
            # `hamevo.evolve_lindblad` builds exp(t·𝓛) on those 2k qubits
            out.append(Op(OpKind.HAMEVO, tuple(op.targets) + tuple(q + n for q in op.targets), (), param=op.param,
                          generator=op.generator, label="HamEvo[Lindblad]", tsym=op.tsym, steps=op.steps, jumps=op.jumps))
            continue
        out.extend(_lift_op(op, n2, 0, False, needs))
        out.extend(_lift_op(op, n2, n, True, needs))
        if op.noise:
            if op.kind in (OpKind.SCALE, OpKind.ADD, OpKind.HAMEVO) or not op.targets:
                raise NotImplementedError(f"digital noise on a {op.label or op.kind.name} op is not supported")
            q = op.targets[0]  # pyqtorch: the gate's (first) target qubit
            out.append(Op(OpKind.MATK, (q, q + n), (), matrix=noise_superoperator(op.noise), label="Noise"))
    return Program(n2, out), sorted(needs)


def extend_values(values: Mapping[str, Any] | None, needs: Sequence[str]) -> dict:
    vals = dict((values or {}).items())  # a ParamTable materialises its lazy entries here
    for nm in needs:
        if nm.startswith(NEGCONJ):
            vals[nm] = -torch.as_tensor(vals[nm[len(NEGCONJ):]]).conj()
        else:
            vals[NEG + nm] = -torch.as_tensor(vals[nm])
    return vals


# ---------------------------------------------------------------------------------------------------
# execution
# ---------------------------------------------------------------------------------------------------
class CompiledDensityProgram:
    """The vectorised form of a (noisy) program, compiled for the engine."""

    def __init__(self, program: Program):
        self.program = program
        self.n_qubits = program.n_qubits
        self.vprogram, self.needs = vectorise(program)
        self.compiled = engine.CompiledProgram(self.vprogram)

    def values(self, values: Mapping[str, Any] | None) -> dict:
        return extend_values(values, self.needs)


def vectorise_state(state: Tensor | None, n: int) -> Tensor | None:
    """`[B, 2^n]` pure states or `[B, 2^n, 2^n]` density matrices → `[B, 4^n]`."""
    if state is None:
        return None
    N = 2**n
    if state.dim() == 3:
        if state.shape[1] != N or state.shape[2] != N:
            raise ValueError(f"density matrix must have shape (batch, 2**{n}, 2**{n}); got {tuple(state.shape)}")
        return state.as_subclass(Tensor).reshape(state.shape[0], N * N)
    if state.dim() != 2 or state.shape[1] != N:
        raise ValueError(f"state must have shape (batch, 2**{n}); got {tuple(state.shape)}")
    return torch.einsum("bi,bj->bij", state, state.conj()).reshape(state.shape[0], N * N)


def run(cd: CompiledDensityProgram, values: Mapping[str, Any] | None = None, state: Tensor | None = None,
        dtype: torch.dtype = torch.complex128, device: torch.device | str | None = None) -> Tensor:
    """ρ_out `[B, 2^n, 2^n]` (a `DensityMatrix`) — `Backend.run` of a circuit whose gates carry digital noise."""
    N = 2**cd.n_qubits
    v = engine.run(cd.compiled, cd.values(values), vectorise_state(state, cd.n_qubits), dtype=dtype, device=device)
    return v.reshape(v.shape[0], N, N).as_subclass(DensityMatrix)


def lift_observable(obs: PauliSum | Program, n: int) -> PauliSum | Program:
    """O → O ⊗ 1 on the doubled register (O acts on the row index of vec(ρ))."""
    if isinstance(obs, PauliSum):
        return _lift_paulisum(obs, 2 * n, 0, False)
    return lift_program(obs, 2 * n, 0, False, set())


def expectation(cd: CompiledDensityProgram, observables: Sequence[Any], values: Mapping[str, Any] | None = None,
                state: Tensor | None = None, dtype: torch.dtype = torch.complex128,
                device: torch.device | str | None = None) -> Tensor:
    """`[B, n_obs]` real `Tr(O ρ)`; `observables` are `engine.CompiledObservable`s of the LIFTED (Hermitian) observables.
    Differentiable w.r.t. the gate parameters by the adjoint walk (see below) and, of course, by `diff_mode="gpsr"` (the
    shift rules hold for noisy expectations); observable coefficients are not differentiated on this path."""
    N = 2**cd.n_qubits
    vals = cd.values(values)
    real_dt = torch.float64 if dtype == torch.complex128 else torch.float32
    needs_grad = torch.is_grad_enabled() and (any(isinstance(x, Tensor) and x.requires_grad for x in vals.values())
                                              or (state is not None and state.requires_grad))
    if needs_grad and cd.compiled.invertible:
        # Tr(Oρ) = ⟨⟨vec(O)|vec(ρ)⟩⟩ for Hermitian O, vec(O) = (O ⊗ 1)|1⟩⟩: a LINEAR functional of the differentiable
        # wavefunction `engine.run` returns for the vectorised program — its backward pass is the adjoint walk with
        # λ = vec(O) (noise superoperators are un-applied with their inverse), so noisy expectations differentiate like
        # noise-free ones in `ad` / `adjoint` mode
        v = engine.run(cd.compiled, vals, vectorise_state(state, cd.n_qubits), dtype=dtype, device=device)
        with torch.no_grad():
            bell = torch.eye(N, dtype=v.dtype, device=v.device).reshape(1, N * N)
            detached = {k: (x.detach() if isinstance(x, Tensor) else x) for k, x in vals.items()}
            ovecs = [engine.apply_observable(cob, bell.clone(), detached) for cob in observables]
        return torch.stack([(o.conj() * v).sum(-1).real for o in ovecs], dim=1).to(real_dt)
    if needs_grad:
        logger.warning("b200sv: this noisy circuit contains non-invertible ops (projectors, AddBlocks): its expectation is not "
                       "differentiable by the adjoint walk; use diff_mode='gpsr'")
    with torch.no_grad():
        v = engine.run(cd.compiled, vals, vectorise_state(state, cd.n_qubits), dtype=dtype, device=device)
        cols = []
        for cob in observables:
            w = engine.apply_observable(cob, v, vals)
            cols.append(w.reshape(w.shape[0], N, N).diagonal(dim1=1, dim2=2).sum(-1).real)
        out = torch.stack(cols, dim=1)
    return out.to(real_dt)


# ---------------------------------------------------------------------------------------------------
# readout noise + sampling
# ---------------------------------------------------------------------------------------------------
class ReadoutNoise:
    """Independent (per-qubit symmetric bit flip) or correlated (full confusion matrix) readout error, applied to the
    outcome probabilities before sampling (`convert_readout_noise`, backends/pyqtorch/convert_ops.py:375-393)."""

    def __init__(self, n_qubits: int, error_probability: Any = None, confusion_matrix: Any = None, **_: Any):
        self.n_qubits = n_qubits
        self.confusion_matrix = None if confusion_matrix is None else np.asarray(torch.as_tensor(confusion_matrix).numpy(), dtype=np.float64)
        if self.confusion_matrix is not None:
            if self.confusion_matrix.shape != (2**n_qubits, 2**n_qubits):
                raise ValueError(f"confusion matrix must have shape (2**{n_qubits}, 2**{n_qubits})")
            self.error_probability = None
            return
        p = 0.1 if error_probability is None else error_probability
        p = np.asarray(torch.as_tensor(p, dtype=torch.float64).reshape(-1).numpy())
        if p.size == 1:
            p = np.repeat(p, n_qubits)
        if p.size != n_qubits or (p < 0).any() or (p > 1).any():
            raise ValueError("readout error_probability must be one probability, or one per qubit, in [0, 1]")
        self.error_probability = p

    def program(self) -> Program:
        """The confusion map as an n-qubit program acting on the probability vector (real 2×2 'gates')."""
        if self.confusion_matrix is not None:
            return Program(self.n_qubits, [Op(OpKind.MATK if self.n_qubits > 1 else OpKind.MAT1, tuple(range(self.n_qubits)), (),
                                              matrix=self.confusion_matrix.astype(np.complex128), label="Confusion")])
        ops = [Op(OpKind.MAT1, (q,), (), matrix=np.array([[1 - p, p], [p, 1 - p]], dtype=np.complex128), label="Confusion")
               for q, p in enumerate(self.error_probability) if p != 0]
        return Program(self.n_qubits, ops)

    def apply(self, probs: Tensor, device: torch.device | str | None = None) -> Tensor:
        """`[B, 2^n]` real probabilities → corrupted probabilities."""
        prog = self.program()
        if not prog.ops:
            return probs
        out = engine.run(engine.CompiledProgram(prog), {}, probs.to(torch.complex128), dtype=torch.complex128, device=device)
        return out.real.clamp_min(0.0)


def probabilities(rho_or_psi: Tensor) -> Tensor:
    """`[B, 2^n]` outcome probabilities of `[B, 2^n, 2^n]` density matrices or `[B, 2^n]` pure states."""
    t = rho_or_psi.as_subclass(Tensor)
    if t.dim() == 3:
        return t.diagonal(dim1=1, dim2=2).real.clamp_min(0.0).to(torch.float64)
    return (t.real.to(torch.float64) ** 2 + t.imag.to(torch.float64) ** 2)


def sample_probabilities(probs: Tensor, n_shots: int, uniforms: Tensor | None = None, little_endian: bool = False) -> list[Counter]:
    """Inverse-CDF sampling (the engine's sum-tree sampler) from explicit probabilities."""
    n = int(probs.shape[1]).bit_length() - 1
    amp = probs.clamp_min(0.0).sqrt().to(torch.complex128).contiguous()
    if uniforms is None:
        uniforms = torch.rand(probs.shape[0], n_shots, dtype=torch.float64, device=probs.device)
    idx = engine.sample_indices(amp, uniforms)
    return engine.counts_from_indices(idx, n, little_endian)
