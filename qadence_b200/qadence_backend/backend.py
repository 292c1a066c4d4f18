"""The qadence `Backend` of the B200-native state-vector simulator.

Drop-in for `/root/reference/qadence/backends/pyqtorch/backend.py:82-323`: same dataclass fields,
same method names, argument meaning and error behaviour (`validate_state` TypeError / ValueError,
`assign_parameters` NotImplementedError, split `{"circuit":…, "observables":…}` dicts, pyqtorch-shaped
`[2]*n+[B]` initial states, `Endianness.LITTLE` views, `list[Counter]` samples).
"""

from __future__ import annotations

import dataclasses
from collections import Counter
from dataclasses import dataclass, field
from logging import getLogger
from typing import Any

import sympy
import torch
from torch import Tensor

from qadence.backend import Backend as BackendInterface
from qadence.backend import Converted, ConvertedCircuit, ConvertedObservable
from qadence.backends.utils import infer_batchsize, is_pyq_shape, unpyqify, validate_state
from qadence.blocks import AbstractBlock
from qadence.circuit import QuantumCircuit
from qadence.measurements import Measurements
from qadence.mitigations import Mitigations
from qadence.noise import NoiseHandler
from qadence.transpile import flatten, scale_primitive_blocks_only, set_noise, transpile
from qadence.types import BackendName, Endianness, Engine, ParamDictType

from .. import density, engine
from ..embedding import ParamTable, UnsupportedExpression, compile_expressions, make_embedding_fn
from ..program import OpKind, Program
from .config import Configuration, default_passes
from .convert_ops import convert_block, convert_observable, convert_readout_noise

logger = getLogger(__name__)
# replace mode (B200SV_REPLACE_PYQTORCH=1, qadence_extensions/__init__.py): this backend answers to the name "pyqtorch"
_REPLACE = "b200sv" not in BackendName.list()
_NAME, _ENGINE = ("pyqtorch", "torch") if _REPLACE else ("b200sv", "b200sv")
_DTYPES = {"complex128": torch.complex128, "complex64": torch.complex64}


def dropout_ops(program: Program, prob: float, mode: str, values: Any) -> set[int]:
    """Indices of the ops one forward pass of a quantum-dropout circuit skips (arXiv:2310.04120, what the reference gets
    from `pyq.DropoutQuantumCircuit`, backends/pyqtorch/backend.py:127-141; pyqtorch is absent here, so the three modes are
    restated from its documented behaviour — parity unpinned, like §4.3's noise):
      rotational_dropout     every parametric rotation whose parameter is trainable is dropped with probability `prob`;
      entangling_dropout     every controlled gate is dropped with probability `prob`;
      canonical_fwd_dropout  rotational dropout, and a dropped rotation takes the entangling gates that follow it on its
                             qubit (up to that qubit's next rotation) with it.
    Draws from torch's global generator (`torch.manual_seed` makes a training run reproducible)."""
    rot = (OpKind.RX, OpKind.RY, OpKind.RZ, OpKind.PHASE)

    def trainable(op: Any) -> bool:
        if not isinstance(op.param, str):
            return False
        if isinstance(values, ParamTable):
            return op.param not in values.nograd
        v = values.get(op.param) if hasattr(values, "get") else None
        return bool(isinstance(v, Tensor) and v.requires_grad)

    ops = list(program.ops)
    coins = torch.bernoulli(torch.full((len(ops),), float(prob))).bool().tolist() if ops else []
    drop: set[int] = set()
    if mode == "entangling_dropout":
        return {i for i, op in enumerate(ops) if op.controls and coins[i]}
    for i, op in enumerate(ops):
        if op.kind in rot and not op.controls and trainable(op) and coins[i]:
            drop.add(i)
            if mode == "canonical_fwd_dropout":
                q = op.targets[0]
                for j in range(i + 1, len(ops)):
                    nxt = ops[j]
                    if nxt.kind in rot and not nxt.controls and q in nxt.targets:
                        break
                    if nxt.controls and (q in nxt.targets or q in nxt.controls):
                        drop.add(j)
    return drop


class NativeCircuit:
    """What `ConvertedCircuit.native` holds: the flat program + its compiled (planned) form.

    Duck-types the attributes qadence reads on pyqtorch circuits: `n_qubits`,
    `dtype`, `device`, `to()`, `operations`, `init_state`, `readout_noise`."""

    def __init__(self, program: Program, dtype: torch.dtype, device: str | None, readout_noise: Any = None):
        self.program = program
        self.is_noisy = program.is_noisy  # gates carry digital noise: runs as a density matrix (qadence_b200/density.py)
        self.compiled = None if self.is_noisy else engine.CompiledProgram(program)
        self._density: density.CompiledDensityProgram | None = None
        self.n_qubits = program.n_qubits
        self.dtype = dtype
        self._device = device
        self.readout_noise = readout_noise
        self.dropout_prob = 0.0  # quantum dropout (pyqtorch `DropoutQuantumCircuit`): set by `Backend.circuit`
        self.dropout_mode = "rotational_dropout"
        self._dropped: dict = {}  # dropped-op index tuple -> CompiledProgram (bounded cache)

    def compiled_for_call(self, training: bool, values: Any) -> "engine.CompiledProgram":
        """The program this call executes: the full one, or — quantum dropout, training mode only — one with a fresh random
        subset of its gates removed (`dropout_ops`)."""
        if not training or self.dropout_prob <= 0.0 or self.compiled is None:
            return self.compiled
        drop = dropout_ops(self.program, self.dropout_prob, str(getattr(self.dropout_mode, "value", self.dropout_mode)), values)
        if not drop:
            return self.compiled
        key = tuple(sorted(drop))
        if key not in self._dropped:
            if len(self._dropped) >= 64:
                self._dropped.pop(next(iter(self._dropped)))
            self._dropped[key] = engine.CompiledProgram(Program(self.program.n_qubits, [op for i, op in enumerate(self.program.ops) if i not in drop]))
        return self._dropped[key]

    @property
    def density(self) -> density.CompiledDensityProgram:
        if self._density is None:
            self._density = density.CompiledDensityProgram(self.program)
        return self._density

    @property
    def operations(self) -> list:
        return self.program.ops

    if _REPLACE:  # replace mode only: this object stands where a `pyqtorch.QuantumCircuit` is expected

        @property  # type: ignore[misc]
        def __class__(self) -> type:
            """`QuantumModel.to()` moves the model only `if isinstance(native, pyqtorch.QuantumCircuit)` (qadence/model.py:651);
            isinstance falls back to `__class__`, `type(obj)` (used by copy / pickle) is untouched."""
            try:
                import pyqtorch

                PyQCircuit = pyqtorch.DropoutQuantumCircuit if self.dropout_prob > 0.0 else pyqtorch.QuantumCircuit  # backend.py:127-141
                return PyQCircuit if isinstance(PyQCircuit, type) else NativeCircuit
            except Exception:
                return NativeCircuit

    @property
    def device(self) -> torch.device:
        return torch.device(self._device) if self._device else torch.device("cuda")

    def to(self, *args: Any, **kwargs: Any) -> "NativeCircuit":
        for a in list(args) + list(kwargs.values()):
            if isinstance(a, torch.dtype) and a in (torch.complex64, torch.complex128):
                self.dtype = a
            elif isinstance(a, (str, torch.device)) and torch.device(a).type == "cuda":
                self._device = str(a)
        return self

    def init_state(self, batch_size: int = 1) -> Tensor:
        return engine.zero_state(self.n_qubits, batch_size, self.dtype, engine._dev(self._device))


class NativeObservable:
    def __init__(self, obs: Any, dtype: torch.dtype):
        self.compiled = engine.CompiledObservable(obs)
        self.obs = obs
        self.dtype = dtype
        self._lifted: engine.CompiledObservable | None = None

    @property
    def lifted(self) -> engine.CompiledObservable:
        """O ⊗ 1 on the doubled register of a vectorised density matrix."""
        if self._lifted is None:
            self._lifted = engine.CompiledObservable(density.lift_observable(self.obs, self.obs.n_qubits))
        return self._lifted

    def to(self, *args: Any, **kwargs: Any) -> "NativeObservable":
        for a in list(args) + list(kwargs.values()):
            if isinstance(a, torch.dtype) and a in (torch.complex64, torch.complex128):
                self.dtype = a
        return self


def _split(param_values: ParamDictType) -> tuple[dict, dict]:
    """`{"circuit":…, "observables":…}` or a flat dict (pyqtorch/backend.py:192-195)."""
    pc = param_values["circuit"] if "circuit" in param_values else param_values
    po = param_values["observables"] if "observables" in param_values else param_values
    return pc, po


def _is_density(state: Any) -> bool:
    return isinstance(state, Tensor) and (isinstance(state, density.DensityMatrix) or type(state).__name__ == "DensityMatrix")


def _reverse_density(rho: Tensor) -> Tensor:
    """Little-endian view of `[B, N, N]` density matrices: the qubit order of the row AND the column index is reversed."""
    B, N, _ = rho.shape
    r = engine.bit_reverse(rho.as_subclass(Tensor).reshape(B * N, N).contiguous()).reshape(B, N, N).transpose(1, 2).contiguous()
    r = engine.bit_reverse(r.reshape(B * N, N)).reshape(B, N, N).transpose(1, 2).contiguous()
    return r.as_subclass(density.DensityMatrix)


def _result_device(cfg: Configuration, *dicts: Any) -> torch.device | None:
    if not cfg.results_on_input_device:
        return None
    devs = set()
    for d in dicts:
        if isinstance(d, Tensor):
            devs.add(d.device)
        elif isinstance(d, ParamTable):
            if d.input_device is not None:
                devs.add(d.input_device)
        elif isinstance(d, dict):
            devs |= {v.device for v in d.values() if isinstance(v, Tensor)}
    if any(dv.type == "cuda" for dv in devs):
        return None
    return torch.device("cpu")


def _device_embedding_fn(conv: Converted, config: Configuration) -> Any:
    """Compile the expressions behind every gate / observable parameter of `conv` for the device
    (replaces the per-call host loop of qadence/blocks/embedding.py:118-167; see qadence_b200/embedding.py)."""
    from qadence.blocks.utils import expressions, parameters, uuid_to_expression
    from qadence.parameters import stringify

    blocks = [conv.circuit.abstract.block] + [o.abstract for o in (conv.observable or [])]
    needed: list[str] = list(conv.circuit.native.program.param_names)
    for o in conv.observable or []:
        for nm in o.native.compiled.names:
            if nm not in needed:
                needed.append(nm)
    symbols: dict[str, Any] = {}
    by_name: dict[str, Any] = {}
    for blk in blocks:
        for p in parameters(blk):
            if isinstance(p, sympy.Array):
                continue
            for sym in (p.free_symbols if isinstance(p, sympy.Basic) else ()):
                symbols[sym.name] = sym
        if config._use_gate_params:
            by_name.update(uuid_to_expression(blk))
        else:
            for e in expressions(blk):
                if not isinstance(e, sympy.Array):
                    by_name[stringify(e)] = e
    if not config._use_gate_params:
        for nm, sym in symbols.items():
            by_name.setdefault(nm, sym)
    if any(getattr(sym, "is_time", False) for sym in symbols.values()):
        raise UnsupportedExpression("time-dependent generator")
    named = []
    for nm in needed:
        e = by_name.get(nm)
        if e is None or isinstance(e, sympy.Array):
            raise UnsupportedExpression(f"no scalar expression behind parameter '{nm}'")
        named.append((nm, e))
    # the reference's dict also has an entry for every gate whose parameter is a plain number (embedding.py:147-153): the
    # parameter-shift machinery looks those uuids up (engines/torch/differentiable_expectation.py:208), so they are table
    # rows too (constant DAGs)
    for nm, e in by_name.items():
        if nm not in needed and not isinstance(e, sympy.Array) and not (isinstance(e, sympy.Basic) and e.is_Symbol and not config._use_gate_params):
            named.append((nm, e))
    prog = compile_expressions(named, is_feature=lambda nm: not getattr(symbols.get(nm), "trainable", False))
    prog.slot_symbols = {nm: frozenset(str(s_) for s_ in e.free_symbols) if isinstance(e, sympy.Basic) else frozenset() for nm, e in named}
    # gate-level names of constant generator matrices (HamEvo of a Tensor): the reference's dict carries them too, and its
    # GPSR machinery looks every uuid of the circuit up (engines/torch/differentiable_expectation.py:207)
    matrices = [(nm, stringify(e)) for nm, e in by_name.items() if isinstance(e, sympy.Array)] if config._use_gate_params else []
    return make_embedding_fn(prog, config.device, conv.embedding_fn, passthrough=not config._use_gate_params, matrix_rows=matrices)


@dataclass(frozen=True, eq=True)
class Backend(BackendInterface):
    """B200-native state-vector backend (hand-written sm_100a CUDA kernels behind a C-ABI)."""

    name: BackendName = BackendName(_NAME)
    supports_ad: bool = True
    support_bp: bool = True
    supports_adjoint: bool = True
    is_remote: bool = False
    with_measurements: bool = True
    with_noise: bool = False
    native_endianness: Endianness = Endianness.BIG
    config: Configuration = field(default_factory=Configuration)
    engine: Engine = Engine(_ENGINE)

    # ------------------------------------------------------------------ conversion
    def convert(self, circuit: QuantumCircuit, observable: list[AbstractBlock] | AbstractBlock | None = None) -> Converted:
        """Inherited algorithm (qadence/backend.py:188-253); parameter embedding is torch's.

        `self.engine` names the differentiation engine `backend_factory` imports
        (qadence/backends/api.py:52), while `embedding()` needs the array library (torch)."""
        proxy = dataclasses.replace(self, engine=Engine.TORCH)
        conv = BackendInterface.convert(proxy, circuit, observable)
        if self.config.device_embedding:
            try:
                conv = Converted(conv.circuit, conv.observable, _device_embedding_fn(conv, self.config), conv.params)
            except UnsupportedExpression as e:
                logger.debug(f"b200sv: keeping the host parameter embedding ({e})")
        return conv

    def circuit(self, circuit: QuantumCircuit) -> ConvertedCircuit:
        passes = self.config.transpilation_passes
        if passes is None:
            passes = default_passes(self.config)
        original = circuit
        if len(passes) > 0:
            circuit = transpile(*passes)(circuit)
        if self.config.noise:  # noise set in the configuration (pyqtorch/backend.py:118-127)
            set_noise(circuit, self.config.noise)
        ops = convert_block(circuit.block, n_qubits=circuit.n_qubits, config=self.config)
        readout = convert_readout_noise(circuit.n_qubits, self.config.noise) if self.config.noise else None
        native = NativeCircuit(Program(circuit.n_qubits, ops), _DTYPES[self.config.dtype], self.config.device, readout)
        if self.config.dropout_probability:  # pyqtorch/backend.py:127-141: a DropoutQuantumCircuit instead of a QuantumCircuit
            if native.is_noisy:
                raise NotImplementedError("quantum dropout of a circuit with digital noise is not supported")
            native.dropout_prob = float(self.config.dropout_probability)
            native.dropout_mode = self.config.dropout_mode or "rotational_dropout"
        return ConvertedCircuit(native=native, abstract=circuit, original=original)

    def _set_block_and_readout_noises(self, circuit: ConvertedCircuit, noise: NoiseHandler | None) -> None:
        """`noise` handed to expectation / sample: set it on the abstract blocks, re-convert the native circuit IN PLACE and
        set the readout error — sticky, like the reference (pyqtorch/backend.py:42-79; pinned by
        tests/qadence/test_noise/test_digital_noise.py:118-127)."""
        if not noise:
            return
        set_noise(circuit, noise)
        old: NativeCircuit = circuit.native
        ops = convert_block(circuit.abstract.block, n_qubits=old.n_qubits, config=self.config)
        readout = convert_readout_noise(circuit.abstract.n_qubits, noise) or old.readout_noise
        circuit.native = NativeCircuit(Program(old.n_qubits, ops), old.dtype, old._device, readout)

    def observable(self, observable: AbstractBlock, n_qubits: int) -> ConvertedObservable:
        block = transpile(flatten, scale_primitive_blocks_only)(observable)  # type: ignore[call-overload]
        native = NativeObservable(convert_observable(block, n_qubits, self.config), _DTYPES[self.config.dtype])
        return ConvertedObservable(native=native, abstract=block, original=observable)

    # ------------------------------------------------------------------ execution
    def _state_in(self, state: Tensor | None, n_qubits: int) -> Tensor | None:
        if state is None:
            return None
        if _is_density(state):
            if state.dim() != 3 or state.shape[1] != 2**n_qubits or state.shape[2] != 2**n_qubits:
                raise ValueError("The initial state must be composed of tensors/arrays of size "
                                 f"(batch_size, 2**n_qubits, 2**n_qubits). Found: {state.shape = }.")  # backends/utils.py:119-129
            return state
        validate_state(state, n_qubits)
        if is_pyq_shape(state, n_qubits) and not (state.dim() == 2 and state.shape[1] == 2**n_qubits):
            state = unpyqify(state)
        return state

    def run(
        self,
        circuit: ConvertedCircuit,
        param_values: dict[str, Tensor] = {},
        state: Tensor | None = None,
        endianness: Endianness = Endianness.BIG,
        pyqify_state: bool = True,
        unpyqify_state: bool = True,
    ) -> Tensor:
        n = circuit.abstract.n_qubits
        nat: NativeCircuit = circuit.native
        state = self._state_in(state, n)
        if nat.is_noisy or _is_density(state):  # mixed state: `[B, 2^n, 2^n]` DensityMatrix (backends/utils.py:140-147)
            out = density.run(nat.density, param_values, state, dtype=nat.dtype, device=nat._device)
            if endianness != self.native_endianness:
                out = _reverse_density(out)
        else:
            out = engine.run(nat.compiled_for_call(circuit.training, param_values), param_values, state, dtype=nat.dtype, device=nat._device)
            if endianness != self.native_endianness:
                if out.requires_grad:  # keep the autograd edge of a differentiable wavefunction: a torch gather, not the kernel
                    idx = torch.tensor([int(format(i, f"0{n}b")[::-1], 2) for i in range(2**n)], device=out.device)
                    out = out.index_select(1, idx)
                else:
                    out = engine.bit_reverse(out)
        dev = _result_device(self.config, param_values, state)
        return out if dev is None else out.to(dev)

    def expectation(
        self,
        circuit: ConvertedCircuit,
        observable: list[ConvertedObservable] | ConvertedObservable,
        param_values: ParamDictType = {},
        state: Tensor | None = None,
        measurement: Measurements | None = None,
        noise: NoiseHandler | None = None,
        mitigation: Mitigations | None = None,
        endianness: Endianness = Endianness.BIG,
    ) -> Tensor:
        self._set_block_and_readout_noises(circuit, noise)  # pyqtorch/backend.py:191
        pc, po = _split(param_values)
        n = circuit.abstract.n_qubits
        nat: NativeCircuit = circuit.native
        observable = observable if isinstance(observable, list) else [observable]
        values = pc if isinstance(pc, ParamTable) and po is pc else dict(pc)
        if po is not pc:
            values.update(po)
        state = self._state_in(state, n)
        if nat.is_noisy or _is_density(state):  # Tr(O ρ); differentiate with diff_mode="gpsr"
            if endianness != self.native_endianness and state is not None:
                raise NotImplementedError("little-endian initial density matrices are not supported")
            out = density.expectation(nat.density, [o.native.lifted for o in observable], values, state, dtype=nat.dtype, device=nat._device)
        elif self.config.shard == "qubits":
            from qadence_b200 import dist as qdist

            if state is not None:
                raise NotImplementedError("a qubit-sharded register starts from |0...0>")
            obs = [o.native.compiled.obs for o in observable]
            if any(o.native.compiled.pauli is None for o in observable):
                raise NotImplementedError("a qubit-sharded register takes Pauli-sum observables")
            out = qdist.sharded_expectation(nat.compiled.program, obs, values, dtype=nat.dtype, device=nat._device)
        else:
            out = engine.expectation(nat.compiled_for_call(circuit.training, values), [o.native.compiled for o in observable], values, state,
                                     dtype=nat.dtype, device=nat._device)
        dev = _result_device(self.config, pc, po, state)
        return out if dev is None else out.to(dev)

    def sample(
        self,
        circuit: ConvertedCircuit,
        param_values: ParamDictType = {},
        n_shots: int = 1,
        state: Tensor | None = None,
        noise: NoiseHandler | None = None,
        mitigation: Mitigations | None = None,
        endianness: Endianness = Endianness.BIG,
        pyqify_state: bool = True,
    ) -> list[Counter]:
        self._set_block_and_readout_noises(circuit, noise)  # pyqtorch/backend.py:298
        n = circuit.abstract.n_qubits
        nat: NativeCircuit = circuit.native
        state = self._state_in(state, n)
        little = endianness != Endianness.BIG
        if nat.is_noisy or nat.readout_noise is not None or _is_density(state):
            # outcome probabilities of ρ (or ψ), corrupted by the readout confusion map, then the same inverse-CDF sampler
            if nat.is_noisy or _is_density(state):
                probs = density.probabilities(density.run(nat.density, param_values, state, dtype=nat.dtype, device=nat._device))
            else:
                probs = density.probabilities(engine.run(nat.compiled, param_values, state, dtype=nat.dtype, device=nat._device))
            if nat.readout_noise is not None:
                probs = nat.readout_noise.apply(probs, device=nat._device)
            samples = density.sample_probabilities(probs, n_shots, little_endian=little)
        else:
            samples = engine.sample(nat.compiled, param_values, n_shots, state, dtype=nat.dtype, device=nat._device, little_endian=little)
        if mitigation is not None:
            from qadence.mitigations.protocols import apply_mitigation

            logger.warning("Mitigation protocol is deprecated. Use qadence-protocols instead.")
            assert noise
            samples = apply_mitigation(noise=noise, mitigation=mitigation, samples=samples)
        return samples

    def assign_parameters(self, circuit: ConvertedCircuit, param_values: ParamDictType) -> Any:
        raise NotImplementedError

    @staticmethod
    def _overlap(bras: Tensor, kets: Tensor) -> Tensor:
        from qadence.overlap import overlap_exact

        return overlap_exact(bras, kets)

    @staticmethod
    def default_configuration() -> Configuration:
        return Configuration()
