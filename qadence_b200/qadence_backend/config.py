"""`Configuration` of the b200sv backend — mirrors
`/root/reference/qadence/backends/pyqtorch/config.py:21-90` (same option names so that user
configuration dicts keep working; options that only make sense for pyqtorch are accepted and
ignored)."""

from __future__ import annotations

from dataclasses import dataclass
from typing import Callable

from qadence.analog import add_background_hamiltonian
from qadence.backend import BackendConfiguration
from qadence.transpile import blockfn_to_circfn, flatten, scale_primitive_blocks_only


def default_passes(config: "Configuration") -> list[Callable]:
    """AnalogBlocks → HamEvo, flatten, push scales to the leaves (pyqtorch/config.py:21-37).

    `chain_single_qubit_ops` is not needed: the fusion planner merges 1-qubit chains itself."""
    return [
        add_background_hamiltonian,
        lambda circ: blockfn_to_circfn(flatten)(circ),
        blockfn_to_circfn(scale_primitive_blocks_only),
    ]


@dataclass
class Configuration(BackendConfiguration):
    # ---- options shared with the pyqtorch backend (same names / defaults) ----
    algo_hevo: str = "krylov"
    """HamEvo algorithm.  "krylov" (default; the reference's "EXP" / "EIG" / "RK4" are accepted and mean the same exact
    propagator): diagonal generators are exponentiated on the fly, everything else uses the matrix-free Lanczos–Krylov
    propagator converged to 1e-13 (the reference's `EXP` is a dense `matrix_exp`; results agree to the parity tolerance).
    "trotter": second-order (Strang) splitting of a Pauli-sum generator into its diagonal part (Rydberg interaction,
    detunings — `b200sv_diag_evolve`) and its single-qubit drive terms (one fused sweep of rotations per step),
    `n_steps_hevo` steps; error O(t·dt²), i.e. NOT within the 1e-10 parity tolerance — an explicit speed / accuracy knob for
    neutral-atom programs whose generator norm makes the Krylov space large."""
    ode_solver: object = None
    """Accepted for compatibility (pyqtorch `SolverType`): time-dependent generators always use the exponential-midpoint
    propagator with `n_steps_hevo` steps."""
    n_steps_hevo: int = 100
    use_gradient_checkpointing: bool = False
    """Accepted for compatibility; the adjoint pass keeps two states, there is nothing to checkpoint."""
    use_single_qubit_composition: bool = False
    """Accepted for compatibility; the fusion planner always merges chains of single-qubit gates."""
    loop_expectation: bool = False
    """Accepted for compatibility; the batched path never materialises per-term states."""
    noise: object = None
    """`NoiseHandler` applied at conversion time (pyqtorch/config.py:62): digital noise on every primitive block — the
    circuit then runs as a density matrix (qadence_b200/density.py) — and/or readout noise for `sample`."""
    dropout_probability: float = 0.0
    """Quantum dropout probability (0 means no dropout), as in pyqtorch/config.py:69-72: while the converted circuit is in
    training mode every `run` / `expectation` executes a fresh random sub-circuit (`qadence_backend.backend.dropout_ops`)."""
    dropout_mode: object = None
    """`DropoutMode` / its string value: "rotational_dropout" (default), "entangling_dropout", "canonical_fwd_dropout"."""
    n_eqs: int | None = None
    shift_prefac: float = 0.5
    gap_step: float = 1.0
    lb: float | None = None
    ub: float | None = None
    # ---- b200sv options ----
    device: str | None = None
    """CUDA device of the simulation (default: current device)."""
    dtype: str = "complex128"
    """State precision: "complex128" (default, qadence/__init__.py:13-15) or "complex64"."""
    device_embedding: bool = True
    """Evaluate the parameter expressions (feature maps, scaled / shared parameters) on the GPU with one kernel launch
    per call instead of the host loop of qadence/blocks/embedding.py; falls back to the host embedding for
    expressions outside the reference's SYMPY_TO_PYQ_MAPPING function set."""
    shard: str = "batch"
    """Multi-GPU axis (one process per GPU, `torch.distributed` initialised by the caller): "batch" — every process runs its
    own slice of the parameter / feature batch on its GPU (nothing to configure: slice the inputs, all-reduce shared
    gradients, `qadence_b200.dist.shard_batch / allreduce_shared_gradients`); "qubits" — ONE register sharded by its
    highest-order qubits over all GPUs of the group (`qadence_b200.dist.PeerShardedSimulator`: sweeps over NVLink peer
    memory, adjoint gradients and Pauli-sum observables included; batch size 1, circuits of simple gates)."""
    results_on_input_device: bool = True
    """Return wavefunctions / expectations on the device of the parameter tensors (CPU for a default
    `QuantumModel`), so the backend is a drop-in even though it always computes on the GPU."""
