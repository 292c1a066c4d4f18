"""Flat op-program IR for the b200sv state-vector backend.

This is what crosses the drop-in boundary instead of pyqtorch's nn.Module tree
(`qadence/backends/pyqtorch/convert_ops.py:177-351` builds `pyq.Sequence/Add/Scale/Merge/...`
modules; we build a flat list of :class:`Op` records).  The same IR is consumed by

* the CUDA engine (`qadence_b200/engine.py`, via the fusion planner `plan.py`), and
* the CPU oracle (`oracle/`, test infrastructure only),

and it is JSON-serialisable so that golden fixtures generated from the reference's own
front-end (`tests/golden/make_golden.py`) can travel to the GPU box where qadence is absent.

Conventions (pinned by `/root/reference/tests/backends/test_endianness.py:113-138`):
qubit 0 is the MOST significant bit of the basis index; states are `[batch, 2**n]` complex.
"""

from __future__ import annotations

import json
from dataclasses import dataclass, field
from enum import IntEnum
from typing import Any, Iterable, Sequence, Union

import numpy as np


class OpKind(IntEnum):
    # --- fusable "simple" gates (applied in registers by the sweep kernel) ---
    MAT1 = 0  # constant 2x2 matrix on one target (X,Y,Z,H,S,T,N,Zero,I,... and their controlled forms)
    RX = 1  # exp(-i θ X/2)   block_to_tensor.py:107-126
    RY = 2  # exp(-i θ Y/2)
    RZ = 3  # exp(-i θ Z/2)
    PHASE = 4  # diag(1, e^{iθ}) block_to_tensor.py:147-157
    SWAP = 5  # swap(t0, t1), optionally controlled (CSWAP)
    SCALE = 6  # ψ ← s·ψ          block_to_tensor.py:455-459
    # --- non-fusable ---
    MATK = 10  # constant dense 2^k x 2^k matrix on k targets (MatrixBlock, Projector)
    ADD = 11  # ψ ← Σ_j sub_j(ψ)  block_to_tensor.py:397-407
    HAMEVO = 12  # ψ ← exp(-i t G) ψ block_to_tensor.py:409-440


# A parameter is either a python number (baked constant) or a string key into the
# embedded `param_values` dict handed to run/expectation (qadence/backend.py:89-105).
Param = Union[float, complex, str]


@dataclass
class PauliTerm:
    """coef * Π_q P_q with P in {X, Y, Z}; `coef` = const * Π values[name] for name in names."""

    const: complex
    names: tuple[str, ...]
    paulis: tuple[tuple[int, str], ...]  # sorted ((qubit, 'X'|'Y'|'Z'), ...)
    expr: str = ""  # time-dependent generators only: symbolic factor of the term, "sin(0.7*t)*x|t,x" (see parse_td_expr)

    def to_json(self) -> dict:
        d = {
            "const": [float(np.real(self.const)), float(np.imag(self.const))],
            "names": list(self.names),
            "paulis": [[q, p] for q, p in self.paulis],
        }
        if self.expr:
            d["expr"] = self.expr
        return d

    @staticmethod
    def from_json(d: dict) -> "PauliTerm":
        return PauliTerm(
            complex(d["const"][0], d["const"][1]),
            tuple(d["names"]),
            tuple((int(q), str(p)) for q, p in d["paulis"]),
            d.get("expr", ""),
        )


@dataclass
class PauliSum:
    """Σ_t coef_t · (Pauli string)_t on `n_qubits` — observables and Pauli-like HamEvo generators."""

    n_qubits: int
    terms: list[PauliTerm] = field(default_factory=list)

    @property
    def param_names(self) -> list[str]:
        out: list[str] = []
        for t in self.terms:
            for nm in t.names:
                if nm not in out:
                    out.append(nm)
        return out

    @property
    def is_diagonal(self) -> bool:
        return all(p == "Z" for t in self.terms for _, p in t.paulis)

    def simplified(self) -> "PauliSum":
        """Merge terms with identical (names, paulis); drop exact zeros."""
        acc: dict[tuple, complex] = {}
        order: list[tuple] = []
        for t in self.terms:
            key = (tuple(sorted(t.names)), t.paulis, t.expr)
            if key not in acc:
                acc[key] = 0j
                order.append(key)
            acc[key] += t.const
        terms = [PauliTerm(acc[k], k[0], k[1], k[2]) for k in order if acc[k] != 0]
        return PauliSum(self.n_qubits, terms)

    @property
    def is_time_dependent(self) -> bool:
        return any(t.expr for t in self.terms)

    def to_json(self) -> dict:
        return {"n_qubits": self.n_qubits, "terms": [t.to_json() for t in self.terms]}

    @staticmethod
    def from_json(d: dict) -> "PauliSum":
        return PauliSum(int(d["n_qubits"]), [PauliTerm.from_json(t) for t in d["terms"]])


@dataclass
class Op:
    kind: OpKind
    targets: tuple[int, ...] = ()
    controls: tuple[int, ...] = ()  # all-ones controls, block_to_tensor.py:175-204
    param: Param | None = None  # angle / scale / evolution time
    matrix: np.ndarray | None = None  # MAT1: [2,2]; MATK: [2^k,2^k] (row-major, qubit targets[0] = MSB)
    subs: list["Program"] | None = None  # ADD
    generator: Any = None  # HAMEVO: PauliSum | Program | np.ndarray [2^k,2^k] | str (key of a matrix value)
    label: str = ""  # reference op name, for diagnostics only
    tsym: str = ""  # HAMEVO with a time-dependent generator: name of the time symbol; `param` is then the DURATION
    steps: int = 0  # ... and the number of integration steps (Configuration.n_steps_hevo); with an empty `tsym`: the number of
    #                 second-order Trotter steps of a constant generator (Configuration.algo_hevo = "trotter"), 0 = exact
    # digital noise channels applied right after this op, in order: ((protocol name, (probabilities...)), ...) — what the
    # reference hands to pyqtorch gates as `noise=` (backends/pyqtorch/convert_ops.py:206-208,354-372); see density.py
    noise: tuple = ()
    # HAMEVO only: Lindblad jump operators L_k, dense [2^k, 2^k] matrices on `targets` (targets[0] = MSB) — the reference's
    # `HamEvo(..., noise_operators=[...])` (backends/pyqtorch/convert_ops.py:249-267); the op then evolves a density matrix
    # under dρ/dt = −i[H, ρ] + Σ_k (L_k ρ L_k† − ½{L_k†L_k, ρ}) (density.py / hamevo.evolve_lindblad)
    jumps: tuple = ()

    def to_json(self) -> dict:
        d: dict[str, Any] = {"kind": int(self.kind), "t": list(self.targets), "c": list(self.controls)}
        if self.label:
            d["label"] = self.label
        if self.noise:
            d["noise"] = [[str(nm), [float(x) for x in pr]] for nm, pr in self.noise]
        if self.jumps:
            d["jumps"] = [{"shape": list(np.shape(m)), "re": np.real(m).ravel().tolist(), "im": np.imag(m).ravel().tolist()} for m in self.jumps]
        if self.tsym:
            d["tsym"] = self.tsym
        if self.steps:
            d["steps"] = self.steps
        if self.param is not None:
            if isinstance(self.param, str):
                d["p"] = self.param
            else:
                c = complex(self.param)
                d["pc"] = [c.real, c.imag]
        if self.matrix is not None:
            m = np.asarray(self.matrix, dtype=np.complex128)
            d["m"] = {"shape": list(m.shape), "re": m.real.ravel().tolist(), "im": m.imag.ravel().tolist()}
        if self.subs is not None:
            d["subs"] = [s.to_json() for s in self.subs]
        if self.generator is not None:
            g = self.generator
            if isinstance(g, PauliSum):
                d["gen"] = {"pauli": g.to_json()}
            elif isinstance(g, Program):
                d["gen"] = {"prog": g.to_json()}
            elif isinstance(g, str):
                d["gen"] = {"key": g}
            else:
                m = np.asarray(g, dtype=np.complex128)
                d["gen"] = {"dense": {"shape": list(m.shape), "re": m.real.ravel().tolist(), "im": m.imag.ravel().tolist()}}
        return d

    @staticmethod
    def from_json(d: dict) -> "Op":
        def mat(x: dict) -> np.ndarray:
            return (np.array(x["re"]) + 1j * np.array(x["im"])).reshape(x["shape"])

        param: Param | None = None
        if "p" in d:
            param = d["p"]
        elif "pc" in d:
            param = complex(d["pc"][0], d["pc"][1])
            if param.imag == 0:
                param = param.real
        gen: Any = None
        if "gen" in d:
            g = d["gen"]
            if "pauli" in g:
                gen = PauliSum.from_json(g["pauli"])
            elif "prog" in g:
                gen = Program.from_json(g["prog"])
            elif "key" in g:
                gen = g["key"]
            else:
                gen = mat(g["dense"])
        return Op(
            kind=OpKind(d["kind"]),
            targets=tuple(d["t"]),
            controls=tuple(d["c"]),
            param=param,
            matrix=mat(d["m"]) if "m" in d else None,
            subs=[Program.from_json(s) for s in d["subs"]] if "subs" in d else None,
            generator=gen,
            label=d.get("label", ""),
            tsym=d.get("tsym", ""),
            steps=int(d.get("steps", 0)),
            noise=tuple((str(nm), tuple(float(x) for x in pr)) for nm, pr in d.get("noise", [])),
            jumps=tuple(mat(x) for x in d.get("jumps", [])),
        )


@dataclass
class Program:
    """An ordered list of ops on `n_qubits`; op k+1 acts on the output of op k."""

    n_qubits: int
    ops: list[Op] = field(default_factory=list)

    # ---- introspection -------------------------------------------------
    def iter_ops(self) -> Iterable[Op]:
        for op in self.ops:
            yield op
            if op.subs:
                for s in op.subs:
                    yield from s.iter_ops()
            if isinstance(op.generator, Program):
                yield from op.generator.iter_ops()

    @property
    def param_names(self) -> list[str]:
        """Every string key this program reads from `param_values`, first-use order."""
        out: list[str] = []

        def add(nm: str) -> None:
            if nm not in out:
                out.append(nm)

        for op in self.iter_ops():
            if isinstance(op.param, str):
                add(op.param)
            if isinstance(op.generator, PauliSum):
                for nm in op.generator.param_names:
                    add(nm)
                for nm in time_dependent_symbols(op.generator, op.tsym):
                    add(nm)
            elif isinstance(op.generator, str):
                add(op.generator)
        return out

    @property
    def is_simple(self) -> bool:
        return all(op.kind < OpKind.MATK for op in self.ops)

    @property
    def is_noisy(self) -> bool:
        # (a Lindblad HAMEVO already lifted to vec(ρ) — 2k targets for 2^k-dimensional jump operators — is an ordinary linear op)
        return any(op.noise or (op.jumps and 2 ** len(op.targets) == np.shape(op.jumps[0])[0]) for op in self.iter_ops())

    def n_gates(self) -> int:
        return sum(1 for _ in self.iter_ops())

    # ---- (de)serialisation --------------------------------------------
    def to_json(self) -> dict:
        return {"n_qubits": self.n_qubits, "ops": [op.to_json() for op in self.ops]}

    @staticmethod
    def from_json(d: dict) -> "Program":
        return Program(int(d["n_qubits"]), [Op.from_json(o) for o in d["ops"]])

    def dumps(self) -> str:
        return json.dumps(self.to_json())

    @staticmethod
    def loads(s: str) -> "Program":
        return Program.from_json(json.loads(s))


def parse_td_expr(s: str):
    """`PauliTerm.expr` is "<sympy expression>|<comma-separated symbol names>" (the explicit symbol list keeps names such
    as `E`, `I`, `S`, `N` from being parsed as sympy singletons).  Returns (expression, [symbol names])."""
    import sympy

    body, _, names = s.partition("|")
    syms = [nm for nm in names.split(",") if nm]
    expr = sympy.sympify(body, locals={nm: sympy.Symbol(nm) for nm in syms})
    return expr, sorted(str(x) for x in expr.free_symbols)


def make_td_expr(expr: Any) -> str:
    return f"{expr}|{','.join(sorted(str(x) for x in expr.free_symbols))}"


def time_dependent_symbols(ps: "PauliSum", tsym: str) -> list[str]:
    """Free symbols (other than the time symbol) of a time-dependent generator's coefficient expressions."""
    out: list[str] = []
    for t in ps.terms:
        if t.expr:
            for sname in parse_td_expr(t.expr)[1]:
                if sname != tsym and sname not in out:
                    out.append(sname)
    return out


def hamevo_td(self: "ProgramBuilder", generator: "PauliSum", duration: Param, tsym: str = "t", steps: int = 100) -> "ProgramBuilder":
    return self._add(Op(OpKind.HAMEVO, (), (), param=duration, generator=generator, label="HamEvo(t)", tsym=tsym, steps=steps))


# ---------------------------------------------------------------------------
# Constant matrices (values follow qadence/blocks/block_to_tensor.py:27-56; TDagger is the true
# conjugate of T here — the reference constant TDAGMAT equals TMAT (block_to_tensor.py:35-36) and
# TDagger is excluded from the pyqtorch backend's supported gates, convert_ops.py:64-65).
# ---------------------------------------------------------------------------
_S2 = 1.0 / np.sqrt(2.0)
CONST_MATS: dict[str, np.ndarray] = {
    "I": np.array([[1, 0], [0, 1]], dtype=np.complex128),
    "X": np.array([[0, 1], [1, 0]], dtype=np.complex128),
    "Y": np.array([[0, -1j], [1j, 0]], dtype=np.complex128),
    "Z": np.array([[1, 0], [0, -1]], dtype=np.complex128),
    "H": np.array([[_S2, _S2], [_S2, -_S2]], dtype=np.complex128),
    "S": np.array([[1, 0], [0, 1j]], dtype=np.complex128),
    "SDagger": np.array([[1, 0], [0, -1j]], dtype=np.complex128),
    "T": np.array([[1, 0], [0, np.exp(1j * np.pi / 4)]], dtype=np.complex128),
    "N": np.array([[0, 0], [0, 1]], dtype=np.complex128),
    "Zero": np.zeros((2, 2), dtype=np.complex128),
    "P0": np.array([[1, 0], [0, 0]], dtype=np.complex128),
    "P1": np.array([[0, 0], [0, 1]], dtype=np.complex128),
}


# ---------------------------------------------------------------------------
# Standalone builder API (what bench.py / smoke() / GPU tests use where qadence is absent).
# Gate names and argument order mirror qadence.operations.
# ---------------------------------------------------------------------------
class ProgramBuilder:
    def __init__(self, n_qubits: int):
        self.p = Program(n_qubits)

    def _add(self, op: Op) -> "ProgramBuilder":
        for q in (*op.targets, *op.controls):
            if not 0 <= q < self.p.n_qubits:
                raise ValueError(f"qubit {q} out of range for {self.p.n_qubits} qubits")
        self.p.ops.append(op)
        return self

    def gate(self, name: str, target: int, controls: Sequence[int] = ()) -> "ProgramBuilder":
        return self._add(Op(OpKind.MAT1, (target,), tuple(controls), matrix=CONST_MATS[name].copy(), label=name))

    def X(self, t: int) -> "ProgramBuilder":
        return self.gate("X", t)

    def Y(self, t: int) -> "ProgramBuilder":
        return self.gate("Y", t)

    def Z(self, t: int) -> "ProgramBuilder":
        return self.gate("Z", t)

    def H(self, t: int) -> "ProgramBuilder":
        return self.gate("H", t)

    def S(self, t: int) -> "ProgramBuilder":
        return self.gate("S", t)

    def T(self, t: int) -> "ProgramBuilder":
        return self.gate("T", t)

    def N(self, t: int) -> "ProgramBuilder":
        return self.gate("N", t)

    def CNOT(self, c: int, t: int) -> "ProgramBuilder":
        return self.gate("X", t, (c,))

    def CZ(self, c: int, t: int) -> "ProgramBuilder":
        return self.gate("Z", t, (c,))

    def Toffoli(self, cs: Sequence[int], t: int) -> "ProgramBuilder":
        return self.gate("X", t, tuple(cs))

    def rot(self, kind: OpKind, t: int, param: Param, controls: Sequence[int] = ()) -> "ProgramBuilder":
        return self._add(Op(kind, (t,), tuple(controls), param=param, label=kind.name))

    def RX(self, t: int, p: Param) -> "ProgramBuilder":
        return self.rot(OpKind.RX, t, p)

    def RY(self, t: int, p: Param) -> "ProgramBuilder":
        return self.rot(OpKind.RY, t, p)

    def RZ(self, t: int, p: Param) -> "ProgramBuilder":
        return self.rot(OpKind.RZ, t, p)

    def PHASE(self, t: int, p: Param) -> "ProgramBuilder":
        return self.rot(OpKind.PHASE, t, p)

    def CRX(self, c: int, t: int, p: Param) -> "ProgramBuilder":
        return self.rot(OpKind.RX, t, p, (c,))

    def CRY(self, c: int, t: int, p: Param) -> "ProgramBuilder":
        return self.rot(OpKind.RY, t, p, (c,))

    def CRZ(self, c: int, t: int, p: Param) -> "ProgramBuilder":
        return self.rot(OpKind.RZ, t, p, (c,))

    def CPHASE(self, c: int, t: int, p: Param) -> "ProgramBuilder":
        return self.rot(OpKind.PHASE, t, p, (c,))

    def U(self, t: int, phi: Param, theta: Param, omega: Param) -> "ProgramBuilder":
        # RZ(omega)·RY(theta)·RZ(phi), block_to_tensor.py:129-144
        return self.RZ(t, phi).RY(t, theta).RZ(t, omega)

    def SWAP(self, a: int, b: int, controls: Sequence[int] = ()) -> "ProgramBuilder":
        return self._add(Op(OpKind.SWAP, (a, b), tuple(controls), label="SWAP"))

    def scale(self, s: Param) -> "ProgramBuilder":
        return self._add(Op(OpKind.SCALE, (), (), param=s, label="Scale"))

    def matrix(self, m: np.ndarray, targets: Sequence[int]) -> "ProgramBuilder":
        m = np.asarray(m, dtype=np.complex128)
        if len(targets) == 1:
            return self._add(Op(OpKind.MAT1, tuple(targets), (), matrix=m, label="Matrix"))
        return self._add(Op(OpKind.MATK, tuple(targets), (), matrix=m, label="Matrix"))

    def add(self, subs: Sequence[Program]) -> "ProgramBuilder":
        return self._add(Op(OpKind.ADD, (), (), subs=list(subs), label="Add"))

    def hamevo(self, generator: Any, time: Param, targets: Sequence[int] = ()) -> "ProgramBuilder":
        return self._add(Op(OpKind.HAMEVO, tuple(targets), (), param=time, generator=generator, label="HamEvo"))

    def noise(self, protocol: str, *error_probability: float) -> "ProgramBuilder":
        """Digital noise channel right after the last gate, on its first target (what `set_noise` + the reference's
        `convert_digital_noise` attach to a block, backends/pyqtorch/convert_ops.py:354-372)."""
        op = self.p.ops[-1]
        op.noise = (*op.noise, (str(getattr(protocol, "value", protocol)), tuple(float(x) for x in error_probability)))
        return self

    def build(self) -> Program:
        return self.p


def hea_program(n_qubits: int, depth: int, prefix: str = "theta") -> Program:
    """Standalone equivalent of `qadence.constructors.hea(n, depth)` (digital, RX-RY-RX + CNOT ladder).

    Gate order and parameter naming follow `/root/reference/qadence/constructors/hea.py`
    (`_rotations_digital`, `_entangler` — even pairs then odd pairs); pinned against the real
    constructor by tests/golden (hea fixtures) whenever the reference is importable.
    """
    b = ProgramBuilder(n_qubits)
    k = 0
    for _ in range(depth):
        for kind in (OpKind.RX, OpKind.RY, OpKind.RX):
            for q in range(n_qubits):
                b.rot(kind, q, f"{prefix}_{k}")
                k += 1
        for start in (0, 1):
            for q in range(start, n_qubits - 1, 2):
                b.CNOT(q, q + 1)
    return b.build()


def total_magnetization(n_qubits: int) -> PauliSum:
    """Σ_q Z_q — `qadence.constructors.total_magnetization` (constructors/hamiltonians.py)."""
    return PauliSum(n_qubits, [PauliTerm(1.0 + 0j, (), ((q, "Z"),)) for q in range(n_qubits)])


ProgramBuilder.hamevo_td = hamevo_td  # type: ignore[attr-defined]
