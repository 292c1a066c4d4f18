"""Generate golden fixtures from the reference's OWN front-end and dense oracle (run in the CPU
container, where /root/reference is importable through tests/_shims):

    python tests/golden/make_golden.py

For each case it builds the circuit with qadence's constructors, converts it with the b200sv
`Backend.convert` (qadence's transpile passes + our `convert_block`), embeds the parameters with
qadence's `embedding_fn`, and records

  program / observables (JSON IR), embedded param values, initial state,
  expected final state    = block_to_tensor(block) @ state   (qadence/blocks/block_to_tensor.py:309,
                            the independent dense oracle used by tests/qadence/test_matrices.py:68-72)
  expected expectation    = Re <psi| block_to_tensor(obs) |psi>
  expected gradients      = torch.autograd through the dense oracle w.r.t. the user-level parameters,
                            mapped onto gate-level entries by the CPU oracle's adjoint (checked equal).

The fixtures travel to the GPU box, where qadence is absent; tests/test_gpu_golden.py replays them
through the CUDA path.
"""

from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path[:0] = [str(ROOT), str(ROOT / "tests" / "_shims"), "/root/reference"]

import qadence  # noqa: E402
from qadence import (  # noqa: E402
    CNOT, CPHASE, CRX, CRY, CRZ, CSWAP, CZ, MCPHASE, MCRX, MCRY, MCRZ, MCZ, PHASE, RX, RY, RZ, SWAP, H, HamEvo, I, N, S,
    SDagger, T, Toffoli, U, X, Y, Z, AnalogInteraction, AnalogRX, AnalogRY, AnalogSWAP, FeatureParameter, Parameter,
    QuantumCircuit, Register, VariationalParameter, add, chain, feature_map, hea, kron, qft, total_magnetization,
    zz_hamiltonian,
)
from qadence.analog import add_background_hamiltonian  # noqa: E402
from qadence.backends.api import backend_factory  # noqa: E402
from qadence.blocks.block_to_tensor import block_to_tensor  # noqa: E402
from qadence.operations import Projector  # noqa: E402

import qadence.states as qstates  # noqa: E402


def _product_state(bitstring: str, batch_size: int = 1, **_: object) -> torch.Tensor:
    s = torch.zeros(batch_size, 2 ** len(bitstring), dtype=torch.cdouble)
    s[:, int(bitstring, 2)] = 1.0
    return s


qstates.product_state = _product_state  # block_to_tensor's Projector branch (the stock one runs pyqtorch)

from oracle import statevec as oracle  # noqa: E402

OUT = Path(__file__).resolve().parent / "cases"
rng = np.random.default_rng(2024)


def rand_state(batch: int, n: int) -> torch.Tensor:
    s = rng.normal(size=(batch, 2**n)) + 1j * rng.normal(size=(batch, 2**n))
    s /= np.linalg.norm(s, axis=1, keepdims=True)
    return torch.tensor(s)


def cplx(t: torch.Tensor) -> dict:
    t = t.detach().to(torch.complex128)
    return {"shape": list(t.shape), "re": t.real.reshape(-1).tolist(), "im": t.imag.reshape(-1).tolist()}


def dense(block, user_vals: dict, n: int, batch: int) -> torch.Tensor:
    """block_to_tensor, batched when the dense oracle can broadcast, else one batch element at a time
    (its Scale / HamEvo branches do not broadcast a batched parameter, block_to_tensor.py:416-420,455-459)."""
    try:
        return block_to_tensor(block, values=user_vals, qubit_support=tuple(range(n)))
    except RuntimeError:
        mats = []
        for b in range(batch):
            vb = {k: (v[b : b + 1] if v.dim() == 1 and v.numel() == batch and batch > 1 else v) for k, v in user_vals.items()}
            mats.append(block_to_tensor(block, values=vb, qubit_support=tuple(range(n))))
        return torch.cat(mats, dim=0)


def make_case(name: str, circuit: QuantumCircuit, inputs: dict, observables=None, state_batch: int | None = None, zero_state: bool = False, src: str = "") -> None:
    bk = backend_factory("b200sv", diff_mode="adjoint")
    conv = bk.convert(circuit, observables)
    n = circuit.n_qubits
    batch = max([1] + [v.numel() for v in inputs.values()])
    params = {k: v.detach().clone() for k, v in conv.params.items()}
    embedded = conv.embedding_fn(params, inputs)
    sb = state_batch or batch
    psi0 = _product_state("0" * n, sb) if zero_state else rand_state(sb, n)
    # ---- dense reference (independent of this repo) ----
    lowered = add_background_hamiltonian(circuit).block
    user_vals = {**{k: v for k, v in params.items() if not k.startswith("fix_")}, **inputs}
    user_vals = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in user_vals.items()}
    umat = dense(lowered, user_vals, n, batch)
    psi_in = psi0.expand(batch, -1) if psi0.shape[0] == 1 and batch > 1 else psi0
    want = torch.einsum("bij,bj->bi", umat.expand(psi_in.shape[0], -1, -1) if umat.shape[0] == 1 else umat, psi_in)
    case = {
        "name": name, "src": src, "n_qubits": n, "batch": batch,
        "program": conv.circuit.native.program.to_json(),
        "values": {k: torch.as_tensor(v).detach().reshape(-1).tolist() for k, v in embedded.items() if torch.as_tensor(v).dim() <= 1},
        "matrix_values": {k: cplx(torch.as_tensor(v)) for k, v in embedded.items() if torch.as_tensor(v).dim() > 1},
        "state": cplx(psi0), "expected_state": cplx(want),
    }
    # ---- the CPU oracle must agree with the dense reference before anything is written ----
    det = {k: torch.as_tensor(v).detach() for k, v in embedded.items()}
    got = oracle.run(conv.circuit.native.program, det, psi0)
    err = (got - want.detach()).abs().max().item()
    assert err < 1e-10, (name, "oracle vs block_to_tensor", err)
    if observables is not None:
        obs_list = observables if isinstance(observables, list) else [observables]
        exps = []
        for ob in obs_list:
            om = dense(ob, user_vals, n, batch)
            om = om.expand(want.shape[0], -1, -1) if om.shape[0] == 1 else om
            exps.append(torch.einsum("bi,bij,bj->b", want.conj(), om, want).real)
        expv = torch.stack(exps, dim=1)
        case["observables"] = [o.native.obs.to_json() for o in conv.observable]
        case["observable_kinds"] = [type(o.native.obs).__name__ for o in conv.observable]
        case["expected_expectation"] = expv.detach().tolist()
        got_e = oracle.expectation(conv.circuit.native.program, [o.native.obs for o in conv.observable], det, psi0)
        assert (got_e - expv.detach()).abs().max().item() < 1e-10, (name, "oracle expectation")
        # gradients of the FIRST observable: dense autograd (user level) vs oracle adjoint (gate level)
        from qadence_b200.program import PauliSum

        if isinstance(conv.observable[0].native.obs, PauliSum) and not conv.observable[0].native.obs.param_names:
            e0, gates = oracle.adjoint_expectation_and_grads(conv.circuit.native.program, conv.observable[0].native.obs, det, psi0)
            case["expected_gate_grads"] = {k: v.tolist() for k, v in gates.items()}
            # chain gate-level grads to user-level through qadence's embedding and compare with dense autograd
            leaf_p = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in params.items()}
            leaf_i = {k: v.clone().requires_grad_(True) for k, v in inputs.items()}
            emb = conv.embedding_fn(leaf_p, leaf_i)
            surrogate = sum((emb[k].reshape(-1) * gates[k]).sum() for k in gates if emb[k].requires_grad)
            leaves = [t for t in list(leaf_p.values()) + list(leaf_i.values()) if t.requires_grad]
            names = [k for k, t in list(leaf_p.items()) + list(leaf_i.items()) if t.requires_grad]
            if leaves and torch.is_tensor(surrogate):
                mine = torch.autograd.grad(surrogate, leaves, allow_unused=True)
                ref_leaves = [user_vals[k] for k in names if k in user_vals]
                ref = torch.autograd.grad(expv[:, 0].sum(), ref_leaves, allow_unused=True)
                for k, a, b_ in zip([k for k in names if k in user_vals], mine, ref):
                    if a is None or b_ is None:
                        continue
                    assert (a - b_).abs().max().item() < 1e-9, (name, "grad", k)
    (OUT / f"{name}.json").write_text(json.dumps(case))
    print(f"{name:32s} n={n} batch={batch} ops={len(conv.circuit.native.program.ops)} oracle-vs-dense {err:.1e}")


def main() -> None:
    OUT.mkdir(parents=True, exist_ok=True)
    for f in OUT.glob("*.json"):
        f.unlink()
    torch.manual_seed(0)
    x2 = {"x": torch.tensor(rng.uniform(-0.99, 0.99, size=2))}
    # --- every gate family (tests/qadence/test_matrices.py:582-771, test_quantum_pyq.py:337-515) ---
    singles = [X(0), Y(1), Z(2), H(0), S(1), SDagger(2), T(0), I(1), N(2), RX(0, 0.5), RY(1, "x"), RZ(2, 0.3), PHASE(0, "x"), U(1, 0.25, 0.5, 0.75)]
    make_case("gates_single", QuantumCircuit(3, chain(*singles)), x2, src="test_quantum_pyq.py:180-316")
    twos = [CNOT(0, 1), CZ(1, 2), SWAP(0, 2), CRX(0, 1, 0.5), CRY(1, 2, "x"), CRZ(2, 0, 0.7), CPHASE(0, 2, "x")]
    make_case("gates_two_qubit", QuantumCircuit(3, chain(*twos)), x2, src="test_quantum_pyq.py:394-515")
    multis = [Toffoli((0, 1), 2), MCZ((0, 1), 3), MCRX((0, 2), 1, 0.4), MCRY((1, 3), 0, "x"), MCRZ((0, 1, 2), 3, 0.9), MCPHASE((3, 1), 0, "x"), CSWAP(0, 1, 3)]
    make_case("gates_multi_controlled", QuantumCircuit(4, chain(H(0), H(1), H(2), H(3), *multis)), x2, src="test_matrices.py multi-controlled")
    make_case("scaled_blocks", QuantumCircuit(2, chain(2 * X(0), RZ(1, 0.5) * 3.0, "x" * CRY(0, 1, 0.2))), {"x": torch.tensor([0.37])}, src="test_quantum_pyq.py:656-681 (batch 1: the dense oracle cannot broadcast a batched Scale)")
    cnot_add = kron(Projector("0", "0", 0), I(1)) + kron(Projector("1", "1", 0), X(1))
    make_case("add_block_cnot", QuantumCircuit(2, chain(X(0), H(1), cnot_add)), {}, src="test_quantum_pyq.py:789-802")
    make_case("projector_two_qubit", QuantumCircuit(3, chain(H(0), H(1), H(2), Projector("10", "10", (1, 2)))), {}, src="block_to_tensor.py:492-507 (ket == bra and contiguous support: the dense oracle builds kron(ket, bra.T) = |bra><ket| and _fill_identities :59-104 mis-places non-contiguous blocks)")
    # --- constructors ---
    obs4 = [total_magnetization(4), zz_hamiltonian(4), add(0.5 * X(0), kron(Y(1), Z(3)), 1.2 * N(2))]
    make_case("cfg1_qnn_hea4_depth2_batch16", QuantumCircuit(4, feature_map(4), hea(4, 2)),
              {"phi": torch.tensor(rng.uniform(-0.99, 0.99, size=16))}, observables=obs4[:1], zero_state=True, state_batch=1,
              src="BASELINE.json configs[0]")
    make_case("hea3_depth2_three_observables", QuantumCircuit(4, feature_map(4, param="x"), hea(4, 1)), {"x": torch.tensor(rng.uniform(-0.99, 0.99, size=3))},
              observables=obs4, src="tests/backends/conftest.py:9-16")
    make_case("qft4", QuantumCircuit(4, qft(4)), {}, observables=[total_magnetization(4)], src="test_matrices.py qft")
    make_case("qft5_hea", QuantumCircuit(5, chain(qft(5), hea(5, 1))), {}, observables=[total_magnetization(5)], src="BASELINE.json configs[4] (small)")
    # --- HamEvo ---
    gen = 3.1 * X(0) + 1.2 * Y(0) + 1.1 * Y(1) + 1.9 * X(1) + 2.4 * Z(0) * Z(1)
    make_case("hamevo_block_generator", QuantumCircuit(2, chain(HamEvo(gen, FeatureParameter("t")), hea(2, 1))),
              {"t": torch.tensor(rng.uniform(0.1, 1.5, size=3))}, observables=[total_magnetization(2)], src="test_quantum_pyq.py:733-760")
    gm = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
    gm = torch.tensor(gm + gm.conj().T)
    make_case("hamevo_tensor_generator", QuantumCircuit(3, chain(H(0), HamEvo(gm, FeatureParameter("t"), qubit_support=(1, 2)))),
              {"t": torch.tensor(rng.uniform(0.1, 1.5, size=2))}, observables=[total_magnetization(3)], src="test_matrices.py HamEvo tensor")
    diag_gen = 0.7 * kron(N(0), N(1)) + 1.3 * kron(N(1), N(2)) + 0.4 * Z(0)
    make_case("hamevo_diagonal_generator", QuantumCircuit(3, chain(H(0), H(1), H(2), HamEvo(diag_gen, FeatureParameter("t")))),
              {"t": torch.tensor(rng.uniform(0.1, 3.0, size=4))}, observables=[add(X(0), X(1), X(2))], src="analog/hamiltonian_terms.py:16-45")
    # --- analog (cfg3, small register) ---
    reg = Register.rectangular_lattice(1, 3, spacing=8.0)
    make_case("cfg3_analog_3atoms", QuantumCircuit(reg, chain(AnalogRX("a"), AnalogInteraction("t"), AnalogRY("b"))),
              {"a": torch.tensor(rng.uniform(0, np.pi, size=4)), "t": torch.tensor(rng.uniform(0, 1000, size=4)), "b": torch.tensor(rng.uniform(0, np.pi, size=4))},
              observables=[total_magnetization(3)], zero_state=True, state_batch=1, src="BASELINE.json configs[2] (3 atoms)")
    reg = Register.rectangular_lattice(2, 2, spacing=8.0)
    make_case("cfg3_analog_4atoms", QuantumCircuit(reg, chain(AnalogRX("a"), AnalogInteraction("t"), AnalogRY("b"))),
              {"a": torch.tensor(rng.uniform(0, np.pi, size=2)), "t": torch.tensor(rng.uniform(0, 1000, size=2)), "b": torch.tensor(rng.uniform(0, np.pi, size=2))},
              observables=[total_magnetization(4)], zero_state=True, state_batch=1, src="BASELINE.json configs[2] (4 atoms)")
    make_case("analog_swap", QuantumCircuit(2, chain(AnalogSWAP(0, 1), SWAP(0, 1))), {}, src="test_quantum_pyq.py:805-823")


if __name__ == "__main__":
    main()
