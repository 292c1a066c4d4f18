"""qadence blocks → flat `Program` / `PauliSum` (replaces
`/root/reference/qadence/backends/pyqtorch/convert_ops.py:177-351`, which builds pyqtorch modules).

Same vocabulary, same argument conventions (Appendix A of SURVEY.md), same error for un-lowered
analog blocks.  Parameters follow `extract_parameter` (`convert_ops.py:81-99`): a non-parametric
block bakes its numeric parameter as a constant, a parametric one refers to the embedded
`param_values` entry by `config.get_param_name(block)` (uuid, or stringified expression).
"""

from __future__ import annotations

import itertools
from typing import Any, Sequence

import numpy as np
import sympy
import torch

from ..program import CONST_MATS, Op, OpKind, PauliSum, PauliTerm, Program

_ROT = {"RX": OpKind.RX, "RY": OpKind.RY, "RZ": OpKind.RZ, "PHASE": OpKind.PHASE}
_FIXED_1Q = {"X", "Y", "Z", "H", "I", "S", "SDagger", "T", "N", "Zero"}
_PAULI_BASIS = {
    "I": np.eye(2, dtype=np.complex128),
    "X": CONST_MATS["X"],
    "Y": CONST_MATS["Y"],
    "Z": CONST_MATS["Z"],
}


def supported_gate_names() -> list[str]:
    """Same set as the pyqtorch backend: every OpName except TDagger (`convert_ops.py:64-65`)."""
    from qadence.types import OpName

    return list(set(OpName.list()) - {OpName.TDAGGER})


def _numeric(x: Any) -> float | complex:
    c = complex(sympy.N(x)) if isinstance(x, sympy.Basic) else complex(x)
    return c.real if c.imag == 0 else c


def extract_parameter(block: Any, config: Any) -> float | complex | str:
    if not block.is_parametric:
        return _numeric(block.parameters.parameter)
    return config.get_param_name(block)[0]


def _projector_matrix(ket: str, bra: str) -> np.ndarray:
    """|ket><bra| with qubit_support[0] as the most significant bit (`block_to_tensor.py:492-507`)."""
    d = 2 ** len(ket)
    m = np.zeros((d, d), dtype=np.complex128)
    m[int(ket, 2), int(bra, 2)] = 1.0
    return m


def convert_digital_noise(noise: Any) -> tuple:
    """`NoiseHandler` → ((protocol name, (error probabilities...)), ...) of its DIGITAL part — the information the reference
    packs into `pyq.noise.DigitalNoiseProtocol` (backends/pyqtorch/convert_ops.py:354-372)."""
    from qadence.types import NoiseProtocol

    digital = noise.filter(NoiseProtocol.DIGITAL)
    if digital is None:
        return ()
    out = []
    for proto, option in zip(digital.protocol, digital.options):
        p = option.get("error_probability")
        probs = tuple(float(x) for x in torch.as_tensor(p, dtype=torch.float64).reshape(-1).tolist())
        out.append((str(getattr(proto, "value", proto)), probs))
    return tuple(out)


def convert_readout_noise(n_qubits: int, noise: Any) -> Any:
    """READOUT part of a `NoiseHandler` → `density.ReadoutNoise` (backends/pyqtorch/convert_ops.py:375-393)."""
    from qadence.types import NoiseProtocol

    from ..density import ReadoutNoise

    readout = noise.filter(NoiseProtocol.READOUT)
    if readout is None:
        return None
    options = dict(readout.options[0])
    if readout.protocol[0] == NoiseProtocol.READOUT.INDEPENDENT:
        options.pop("confusion_matrix", None)
    elif options.get("confusion_matrix") is None:
        raise ValueError("correlated readout noise needs a `confusion_matrix` option")
    return ReadoutNoise(n_qubits, **options)


def _with_noise(ops: list[Op], noise: tuple) -> list[Op]:
    if noise:
        ops[-1].noise = noise
    return ops


def convert_block(block: Any, n_qubits: int | None = None, config: Any = None) -> list[Op]:
    """Block tree → ordered list of ops (an `AddBlock` becomes one ADD op holding sub-programs)."""
    from qadence.blocks import (
        AddBlock,
        CompositeBlock,
        MatrixBlock,
        ParametricBlock,
        ScaleBlock,
        TimeEvolutionBlock,
    )
    from qadence.blocks.primitive import ProjectorBlock
    from qadence.operations import (
        U,
        multi_qubit_gateset,
        non_unitary_gateset,
        single_qubit_gateset,
        three_qubit_gateset,
        two_qubit_gateset,
    )

    qs = block.qubit_support
    if n_qubits is None:
        n_qubits = max(qs) + 1
    # digital noise set on a primitive block (`set_noise`) travels on the op and is applied by the density-matrix path
    # (qadence_b200/density.py); like the reference, Scale / HamEvo / composite blocks do not take it (convert_ops.py:206-269)
    noise = convert_digital_noise(block.noise) if getattr(block, "noise", None) else ()

    if isinstance(block, ScaleBlock):
        return [*convert_block(block.block, n_qubits, config), Op(OpKind.SCALE, param=extract_parameter(block, config), label="Scale")]

    if isinstance(block, TimeEvolutionBlock):
        gen = block.generator
        jumps = _convert_jumps(block)
        if getattr(gen, "is_time_dependent", False):
            op = _convert_time_dependent(block, n_qubits, config)
            op.jumps = jumps
            return [op]
        names = config.get_param_name(block)
        if isinstance(gen, sympy.Basic):  # symbolic matrix generator: looked up in param_values (convert_ops.py:231-232)
            generator: Any = names[1]
        elif isinstance(gen, torch.Tensor):
            m = gen.detach().to(torch.complex128).numpy()
            m = m.reshape(m.shape[-2], m.shape[-1])
            if not np.allclose(m, m.conj().T, atol=1e-10):
                raise ValueError("HamEvo generator matrix must be Hermitian")
            generator = m
        else:
            generator = convert_generator(gen, n_qubits, config)
        steps = 0
        if str(getattr(config, "algo_hevo", "")).lower() == "trotter":  # second-order Trotter splitting (hamevo.evolve_trotter)
            if isinstance(generator, PauliSum):
                steps = max(1, int(config.n_steps_hevo))
        if jumps:
            steps = 0  # the Lindblad propagator is always the exact exponential (hamevo.evolve_lindblad)
            if not isinstance(generator, (np.ndarray, str, PauliSum)):
                raise NotImplementedError("b200sv: a HamEvo with `noise_operators` needs a Pauli-like or matrix generator")
        return [Op(OpKind.HAMEVO, tuple(qs), (), param=names[0], generator=generator, label="HamEvo", steps=steps, jumps=jumps)]

    if isinstance(block, MatrixBlock):
        m = block.matrix.detach().to(torch.complex128).numpy()
        m = m.reshape(m.shape[-2], m.shape[-1])
        kind = OpKind.MAT1 if len(qs) == 1 else OpKind.MATK
        return _with_noise([Op(kind, tuple(qs), (), matrix=m, label="MatrixBlock")], noise)

    if isinstance(block, CompositeBlock):
        if isinstance(block, AddBlock):
            return [Op(OpKind.ADD, subs=[Program(n_qubits, convert_block(b, n_qubits, config)) for b in block.blocks], label="Add")]
        return list(itertools.chain.from_iterable(convert_block(b, n_qubits, config) for b in block.blocks))

    if isinstance(block, tuple(non_unitary_gateset)):
        if isinstance(block, ProjectorBlock):
            if block.name == "N":  # N(t, state="0") keeps name "N": |1><1| like the reference (convert_ops.py:284-285)
                return _with_noise([Op(OpKind.MAT1, (q,), (), matrix=CONST_MATS["N"].copy(), label="N") for q in qs], noise)
            m = _projector_matrix(block.ket, block.bra)
            return _with_noise([Op(OpKind.MAT1 if len(qs) == 1 else OpKind.MATK, tuple(qs), (), matrix=m, label="Projector")], noise)
        return [Op(OpKind.MAT1, (qs[0],), (), matrix=CONST_MATS["Zero"].copy(), label="Zero")]

    if isinstance(block, tuple(single_qubit_gateset)):
        if isinstance(block, ParametricBlock):
            if isinstance(block, U):  # RZ(omega)·RY(theta)·RZ(phi), ParamMap order phi, theta, omega
                phi, theta, omega = (
                    config.get_param_name(block) if block.is_parametric else [_numeric(p) for p in block.parameters.expressions()]
                )
                t = qs[0]
                return _with_noise([Op(OpKind.RZ, (t,), (), param=phi, label="U.phi"), Op(OpKind.RY, (t,), (), param=theta, label="U.theta"),
                                    Op(OpKind.RZ, (t,), (), param=omega, label="U.omega")], noise)
            return _with_noise([Op(_ROT[block.name], (qs[0],), (), param=extract_parameter(block, config), label=str(block.name))], noise)
        return _with_noise([Op(OpKind.MAT1, (qs[0],), (), matrix=CONST_MATS[str(block.name)].copy(), label=str(block.name))], noise)

    if isinstance(block, tuple(two_qubit_gateset)) or isinstance(block, tuple(three_qubit_gateset) + tuple(multi_qubit_gateset)):
        name = str(block.name)
        base = name[1:] if name.startswith("M") else name  # MCRX -> CRX, MCZ -> CZ (convert_ops.py:327)
        if base in ("SWAP",):
            return _with_noise([Op(OpKind.SWAP, (qs[0], qs[1]), (), label="SWAP")], noise)
        if base == "CSWAP":
            return _with_noise([Op(OpKind.SWAP, tuple(qs[-2:]), tuple(qs[:-2]), label="CSWAP")], noise)
        controls, target = tuple(qs[:-1]), qs[-1]
        if isinstance(block, ParametricBlock):
            return _with_noise([Op(_ROT[base[1:]], (target,), controls, param=extract_parameter(block, config), label=name)], noise)
        inner = {"CNOT": "X", "CZ": "Z", "Toffoli": "X"}[base]
        return _with_noise([Op(OpKind.MAT1, (target,), controls, matrix=CONST_MATS[inner].copy(), label=name)], noise)

    raise NotImplementedError(
        f"Non supported operation of type {type(block)}. "
        "In case you are trying to run an `AnalogBlock`, make sure you "
        "specify the `device_specs` in your `Register` first."
    )


def _convert_time_dependent(block: Any, n_qubits: int, config: Any) -> Op:
    """HamEvo whose generator carries a TimeParameter: the reference switches to expression-level parameter names,
    takes `duration` from the block and hands the generator to pyqtorch's ODE solvers with `n_steps_hevo` steps
    (backends/pyqtorch/convert_ops.py:227-231,258-268).  Here the generator becomes a Pauli sum whose coefficients keep
    their symbolic time dependence; the engine integrates it with `n_steps_hevo` exponential-midpoint steps."""
    config._use_gate_params = False
    times = {str(s_) for s_ in _scale_symbols(block.generator) if getattr(s_, "is_time", False)}
    if len(times) != 1:
        raise NotImplementedError(f"time-dependent HamEvo needs exactly one TimeParameter, found {sorted(times)}")
    try:
        ps = to_pauli_sum(block.generator, n_qubits, config, symbolic=True)
    except NotPauliLike as e:
        raise NotImplementedError(f"time-dependent HamEvo generator is not a sum of Pauli-like strings ({e})") from e
    for t in ps.terms:
        if abs(complex(t.const).imag) > 1e-12:
            raise ValueError("HamEvo generator must be Hermitian")
    duration = block.duration
    if isinstance(duration, sympy.Basic) and duration.free_symbols:
        dur: Any = config.get_param_name(block)[1]  # ("t", "duration", ...) — convert_ops.py:263
    else:
        dur = float(_numeric(duration))
    return Op(OpKind.HAMEVO, tuple(block.qubit_support), (), param=dur, generator=ps, label="HamEvo(t)",
              tsym=times.pop(), steps=int(config.n_steps_hevo))


def _convert_jumps(block: Any) -> tuple:
    """`HamEvo.noise_operators` → dense jump operators on the HamEvo's qubit support (backends/pyqtorch/convert_ops.py:249-256;
    the constructor, operations/ham_evo.py:157-170, already rejects parametric operators and supports outside the generator's)."""
    ops = getattr(block, "noise_operators", None) or []
    if len(ops) == 0:
        return ()
    from qadence.blocks.block_to_tensor import block_to_tensor

    qs = tuple(block.qubit_support)
    if len(qs) > 5:
        raise NotImplementedError("b200sv: `noise_operators` are supported on HamEvo blocks of up to 5 qubits (dense superoperator)")
    out = []
    for nb in ops:
        m = block_to_tensor(nb, {}, qubit_support=qs, use_full_support=False)
        out.append(m.detach().to(torch.complex128).reshape(m.shape[-2], m.shape[-1]).numpy())
    return tuple(out)


def _scale_symbols(block: Any) -> set:
    from qadence.blocks import CompositeBlock, ScaleBlock

    out: set = set()
    if isinstance(block, ScaleBlock):
        p = block.parameters.parameter
        if isinstance(p, sympy.Basic):
            out |= set(p.free_symbols)
        out |= _scale_symbols(block.block)
    elif isinstance(block, CompositeBlock):
        for b in block.blocks:
            out |= _scale_symbols(b)
    return out


# ---------------------------------------------------------------------------------------------------
# Pauli-sum extraction (observables, HamEvo generators)
# ---------------------------------------------------------------------------------------------------
class NotPauliLike(Exception):
    pass


_EXPR = "\0expr:"  # marker of a symbolic (time-dependent) factor inside a term's `names` tuple


def _terms(block: Any, config: Any, symbolic: bool = False) -> list[tuple[complex, tuple[str, ...], dict[int, np.ndarray]]]:
    """Expand a block into Σ const·Π names · ⊗_q (2x2 matrix on q); raises NotPauliLike otherwise.
    With `symbolic`, non-numeric scales are kept as sympy expressions (time-dependent generators)."""
    from qadence.blocks import AddBlock, ChainBlock, KronBlock, ParametricBlock, PrimitiveBlock, ScaleBlock
    from qadence.blocks.primitive import ProjectorBlock

    if symbolic:
        _t = lambda b: _terms(b, config, True)  # noqa: E731
    else:
        _t = lambda b: _terms(b, config)  # noqa: E731
    if isinstance(block, ScaleBlock) and symbolic:
        inner = _t(block.block)
        p = block.parameters.parameter
        if isinstance(p, sympy.Basic) and p.free_symbols:
            key = str(len(_SYMS))
            _SYMS[key] = p
            return [(c, (*nm, _EXPR + key), f) for c, nm, f in inner]
        return [(c * complex(_numeric(p)), nm, f) for c, nm, f in inner]
    if isinstance(block, ScaleBlock):
        inner = _terms(block.block, config)
        p = extract_parameter(block, config)
        if isinstance(p, str):
            return [(c, (*nm, p), f) for c, nm, f in inner]
        return [(c * complex(p), nm, f) for c, nm, f in inner]
    if isinstance(block, AddBlock):
        return list(itertools.chain.from_iterable(_t(b) for b in block.blocks))
    if isinstance(block, (ChainBlock, KronBlock)):
        acc: list[tuple[complex, tuple[str, ...], dict[int, np.ndarray]]] = [(1.0 + 0j, (), {})]
        for b in block.blocks:  # later blocks act after earlier ones: factor_q <- new_q @ factor_q
            nxt = []
            for c1, n1, f1 in acc:
                for c2, n2, f2 in _t(b):
                    f = dict(f1)
                    for q, mat in f2.items():
                        f[q] = mat @ f[q] if q in f else mat
                    nxt.append((c1 * c2, (*n1, *n2), f))
            acc = nxt
            if len(acc) > 4096:
                raise NotPauliLike("too many terms")
        return acc
    if isinstance(block, ProjectorBlock):
        if block.name == "N":
            return [(1.0 + 0j, (), {q: CONST_MATS["N"] for q in block.qubit_support})]
        fac = {}
        for q, k, b_ in zip(block.qubit_support, block.ket, block.bra):
            m = np.zeros((2, 2), dtype=np.complex128)
            m[int(k), int(b_)] = 1.0
            fac[q] = m
        return [(1.0 + 0j, (), fac)]
    if isinstance(block, PrimitiveBlock) and not isinstance(block, ParametricBlock) and str(block.name) in _FIXED_1Q:
        return [(1.0 + 0j, (), {block.qubit_support[0]: CONST_MATS[str(block.name)]})]
    raise NotPauliLike(type(block).__name__)


_SYMS: dict[str, Any] = {}  # id → sympy expression, filled while a symbolic expansion is in flight


def to_pauli_sum(block: Any, n_qubits: int, config: Any, symbolic: bool = False) -> PauliSum:
    """Σ coef·(Pauli string) in the {I,X,Y,Z} basis; N, projectors, S, T, H ... are expanded."""
    from ..program import make_td_expr

    out: list[PauliTerm] = []
    _SYMS.clear()
    for const, names, factors in _terms(block, config, symbolic):
        expr = ""
        sym = [_SYMS[nm[len(_EXPR):]] for nm in names if nm.startswith(_EXPR)]
        if sym:
            prod = sympy.Integer(1)
            for e_ in sym:
                prod = prod * e_
            num, rest = prod.as_coeff_Mul()
            const = const * complex(num)
            expr = make_td_expr(rest) if rest.free_symbols else ""
            names = tuple(nm for nm in names if not nm.startswith(_EXPR))
        comps: list[list[tuple[str, complex]]] = []
        qubits = sorted(factors)
        for q in qubits:
            m = factors[q]
            cq = [(p, complex(np.trace(_PAULI_BASIS[p] @ m) / 2)) for p in "IXYZ"]
            cq = [(p, c) for p, c in cq if abs(c) > 1e-15]
            if not cq:
                comps = []
                const = 0
                break
            comps.append(cq)
        if const == 0:
            continue
        for combo in itertools.product(*comps) if comps else [()]:
            c = complex(const)
            paulis = []
            for q, (p, cp) in zip(qubits, combo):
                c *= cp
                if p != "I":
                    paulis.append((q, p))
            out.append(PauliTerm(c, tuple(names), tuple(paulis), expr))
    return PauliSum(n_qubits, out).simplified()


def convert_generator(gen: Any, n_qubits: int, config: Any) -> PauliSum | Program:
    try:
        return to_pauli_sum(gen, n_qubits, config)
    except NotPauliLike:
        return Program(n_qubits, convert_block(gen, n_qubits, config))


def convert_observable(block: Any, n_qubits: int, config: Any) -> PauliSum | Program:
    return convert_generator(block, n_qubits, config)
