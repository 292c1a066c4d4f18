"""Randomised differential test of `convert_observable` (CPU, needs /root/reference): random sums of scaled Pauli / N / I
strings with parametric coefficients -> PauliSum -> oracle expectation, against the reference's dense `block_to_tensor`.

    python tools/cpu_fuzz_observable.py [seed] [seconds]
"""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'tests'))
from conftest import enable_reference
enable_reference()
import numpy as np, torch, sympy
import qadence
import qadence.states as qstates
def _product_state(bitstring, batch_size=1, **_):
    s_ = torch.zeros(batch_size, 2 ** len(bitstring), dtype=torch.cdouble); s_[:, int(bitstring, 2)] = 1.0; return s_
qstates.product_state = _product_state
from qadence import QuantumCircuit, chain, kron, add, backend_factory
from qadence.blocks.block_to_tensor import block_to_tensor
from qadence.parameters import VariationalParameter
import qadence.operations as O
from oracle import statevec as oracle
from oracle import embedding as oemb
from qadence_b200 import embedding as emb
from qadence_b200.program import ProgramBuilder
emb.evaluate = oemb.evaluate; emb._resolve_device = lambda d: torch.device("cpu")
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 60
w = VariationalParameter("w")
def term(n):
    qs = [int(q) for q in rng.permutation(n)][: int(rng.integers(1, n + 1))]
    ops = [[O.X, O.Y, O.Z, O.N, O.I][int(rng.integers(5))](q) for q in qs]
    blk = kron(*ops) if len(ops) > 1 else ops[0]
    k = rng.random()
    if k < 0.4: return float(rng.normal()) * blk
    if k < 0.6: return (w if rng.random() < 0.5 else 2 * w + 0.1) * blk
    return blk
t0 = time.time(); ok = 0; skipped = 0
while time.time() - t0 < budget:
    n = int(rng.integers(1, 5))
    obs = add(*[term(n) for _ in range(int(rng.integers(1, 6)))])
    if rng.random() < 0.3: obs = float(rng.normal()) * obs
    bk = backend_factory("b200sv", diff_mode="ad")
    circ = QuantumCircuit(n, chain(*[O.RX(q, 0.3 + q) for q in range(n)], *[O.CNOT(q, q + 1) for q in range(n - 1)], O.I(n - 1)))
    try:
        conv = bk.convert(circ, obs)
    except (NotImplementedError, SyntaxError, ValueError, AssertionError) as e:
        skipped += 1; continue
    vals = conv.embedding_fn(conv.params, {})
    got = oracle.expectation(conv.circuit.native.program, [o.native.compiled.obs for o in conv.observable], {k: torch.as_tensor(v).detach() for k, v in dict(vals).items()})
    user = {k: v for k, v in conv.params.items() if not k.startswith("fix_")}
    U = block_to_tensor(circ.block, values=user, qubit_support=tuple(range(n)))
    psi = U[:, :, 0]
    Om = block_to_tensor(obs, values=user, qubit_support=tuple(range(n)))
    want = torch.einsum("bi,bij,bj->b", psi.conj(), Om, psi).real.detach()
    err = (got.reshape(-1) - want.reshape(-1)).abs().max().item()
    if err > 1e-10:
        print("MISMATCH", n, err, got, want); print(obs); sys.exit(1)
    ok += 1
print("ok:", ok, "skipped", skipped)
