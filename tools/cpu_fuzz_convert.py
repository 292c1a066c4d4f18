"""Randomised differential test of `convert_block` (CPU, needs /root/reference): random qadence blocks (all of Appendix A: fixed and
parametric 1-qubit gates, U, controlled / multi-controlled gates, SWAP / CSWAP / Toffoli, scaled blocks, kron / chain / add)
with parameter expressions -> flat program -> oracle, against the reference's own dense `block_to_tensor`.

    python tools/cpu_fuzz_convert.py [seed] [seconds]
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
from qadence.parameters import FeatureParameter, VariationalParameter
import qadence.operations as O
from oracle import statevec as oracle
from oracle import embedding as oemb
from qadence_b200 import embedding as emb
emb.evaluate = oemb.evaluate; emb._resolve_device = lambda d: torch.device("cpu")
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 60
x = FeatureParameter("x"); w = VariationalParameter("w"); v = VariationalParameter("v")
def rand_param():
    k = rng.integers(0, 6)
    return [float(rng.normal()), x, w, 2 * x + w, sympy.sin(x) * v, w ** 2 + 0.3][int(k)]
def rand_gate(n):
    qs = [int(q) for q in rng.permutation(n)]
    k = int(rng.integers(0, 22))
    fixed1 = [O.X, O.Y, O.Z, O.H, O.S, O.SDagger, O.T, O.I, O.N]
    if k < 6: return fixed1[int(rng.integers(len(fixed1)))](qs[0])
    if k < 10: return [O.RX, O.RY, O.RZ, O.PHASE][k - 6](qs[0], rand_param())
    if k == 10: return O.U(qs[0], rand_param(), rand_param(), rand_param())
    if n >= 2:
        if k == 11: return O.CNOT(qs[0], qs[1])
        if k == 12: return O.CZ(qs[0], qs[1])
        if k == 13: return O.SWAP(qs[0], qs[1])
        if k < 18: return [O.CRX, O.CRY, O.CRZ, O.CPHASE][k - 14](qs[0], qs[1], rand_param())
    if n >= 3:
        if k == 18: return O.Toffoli((qs[0], qs[1]), qs[2])
        if k == 19: return O.CSWAP(qs[0], qs[1], qs[2])
        if k == 20: return O.MCRX((qs[0], qs[1]), qs[2], rand_param())
        if k == 21: return O.MCZ((qs[0], qs[1]), qs[2])
    return O.H(qs[0])
def rand_block(n, depth):
    k = int(rng.integers(0, 10))
    if depth == 0 or k < 4: return rand_gate(n)
    if k < 7: return chain(*[rand_block(n, depth - 1) for _ in range(int(rng.integers(1, 5)))])
    if k == 7:  # kron of gates on disjoint qubits
        qs = [int(q) for q in rng.permutation(n)][: int(rng.integers(1, n + 1))]
        return kron(*[[O.X, O.H, O.Z, O.S][int(rng.integers(4))](q) if rng.random() < 0.5 else O.RX(q, rand_param()) for q in qs])
    if k == 8: return rand_param() * rand_block(n, depth - 1) if rng.random() < 0.5 else float(rng.normal()) * rand_block(n, depth - 1)
    return add(*[rand_block(n, depth - 1) for _ in range(int(rng.integers(1, 4)))])
t0 = time.time(); ok = 0
while time.time() - t0 < budget:
    n = int(rng.integers(1, 5))
    try:
        block = chain(rand_block(n, int(rng.integers(0, 4))), O.I(n - 1))
    except Exception as e:
        continue
    bk = backend_factory("b200sv", diff_mode=None)
    try:
        conv = bk.convert(QuantumCircuit(n, block))
    except (NotImplementedError, SyntaxError) as e:
        continue
    inputs = {"x": torch.tensor([0.37])}
    vals = conv.embedding_fn(conv.params, inputs)
    psi = torch.tensor(rng.normal(size=(1, 2**n)) + 1j * rng.normal(size=(1, 2**n)))
    user = {**{k: v_ for k, v_ in conv.params.items() if not k.startswith("fix_")}, **inputs}
    try:
        dense = block_to_tensor(block, values=user, qubit_support=tuple(range(n)))
    except Exception as e:
        dense_fail = globals().get('dense_fail', 0) + 1; globals()['dense_fail'] = dense_fail; last = repr(e)[:200]; continue
    want = torch.einsum("bij,bj->bi", dense, psi).detach()
    got = oracle.run(conv.circuit.native.program, {k: torch.as_tensor(v_).detach() for k, v_ in dict(vals).items()}, psi)
    err = (got - want).abs().max().item() / max(1.0, want.abs().max().item())
    if err > 1e-10:
        print("MISMATCH", n, err); print(block); sys.exit(1)
    ok += 1
print("ok:", ok, "dense failures", globals().get("dense_fail", 0), globals().get("last"))
