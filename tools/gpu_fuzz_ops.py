"""Randomised differential test of the non-sweep kernels against the CPU oracle: Pauli-sum expectation / application
(random X/Y/Z strings, single-Z terms, identity terms, parametric coefficients), diagonal evolution, Krylov HamEvo,
dense k-qubit matrices, bit-exact sampling on adversarial uniforms, bit reversal.

    python tools/gpu_fuzz_ops.py [seconds] [seed]
"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch

from oracle import statevec as oracle
from qadence_b200 import engine
from qadence_b200.program import Op, OpKind, PauliSum, PauliTerm, Program, ProgramBuilder

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 40.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
dev = torch.device("cuda:0")
t0 = time.time()
count = {"pauli": 0, "hamevo": 0, "dense": 0, "sample": 0}


def rand_state(B, n):
    s = torch.tensor(rng.normal(size=(B, 2**n)) + 1j * rng.normal(size=(B, 2**n)))
    return s / s.abs().pow(2).sum(1, keepdim=True).sqrt()


def rand_paulisum(n, diagonal=False, names=()):
    terms = []
    for _ in range(int(rng.integers(1, 9))):
        k = int(rng.integers(0, min(n, 4) + 1))
        qs = sorted(int(q) for q in rng.choice(n, size=k, replace=False))
        ps = tuple((q, "Z" if diagonal else "XYZ"[int(rng.integers(3))]) for q in qs)
        nm = tuple(nm for nm in names if rng.random() < 0.3)
        terms.append(PauliTerm(float(rng.normal()), nm, ps))
    return PauliSum(n, terms)


def fail(what, *info):
    print("MISMATCH", what, *info)
    sys.exit(1)


while time.time() - t0 < budget:
    n = int(rng.integers(1, 12))
    B = int(rng.integers(1, 4))
    psi = rand_state(B, n)
    vals = {"g": torch.tensor(rng.uniform(0.3, 1.5, size=B)), "h": torch.tensor(rng.uniform(-1, 1, size=1))}
    ident = ProgramBuilder(n).build()
    # ---- Pauli-sum expectation and application ----
    ps = rand_paulisum(n, names=("g", "h"))
    cps = engine.CompiledPauliSum.build(ps)
    coef = cps.coef({k: v.to(dev) for k, v in vals.items()}, dev)
    e = engine.pauli_expectation(cps, coef, psi.to(dev)).cpu()
    want = oracle.expectation(ident, [ps], vals, psi)[:, 0]
    if (e - want).abs().max() > 1e-11:
        fail("pauli_expectation", n, ps.to_json(), (e - want).abs().max().item())
    applied = engine.pauli_apply(cps, coef, psi.to(dev)).cpu()
    wa = oracle.from_pyq_shape(oracle.paulisum_apply(ps.simplified(), oracle.to_pyq_shape(psi, n), vals))
    if (applied - wa).abs().max() > 1e-11:
        fail("pauli_apply", n, ps.to_json(), (applied - wa).abs().max().item())
    count["pauli"] += 1
    # ---- HamEvo: diagonal and Krylov, parametric time ----
    if n <= 9:
        for diagonal in (True, False):
            gen = rand_paulisum(n, diagonal=diagonal, names=("g",))
            prog = ProgramBuilder(n).hamevo(gen, "t").build()
            hv = {**vals, "t": torch.tensor(rng.uniform(0.1, 3.0, size=B))}
            got = engine.run(engine.CompiledProgram(prog), hv, psi, device=dev).cpu()
            wh = oracle.run(prog, hv, psi)
            if (got - wh).abs().max() > 1e-10:
                exact = None
                if diagonal and n == 1:  # closed form for the record
                    c0 = sum(complex(tm.const) * (1 if not tm.paulis else 0) for tm in gen.terms)
                    cz = sum(complex(tm.const) for tm in gen.terms if tm.paulis)
                    ph = torch.stack([torch.exp(-1j * hv["t"] * (c0 + cz)), torch.exp(-1j * hv["t"] * (c0 - cz))], 1)
                    exact = psi * ph
                fail("hamevo", n, diagonal, gen.to_json(), (got - wh).abs().max().item(), "t", hv["t"].tolist(), "g", hv["g"].tolist(),
                     "gpu-vs-closed-form", None if exact is None else (got - exact).abs().max().item(),
                     "oracle-vs-closed-form", None if exact is None else (wh - exact).abs().max().item())
            count["hamevo"] += 1
    # ---- dense k-qubit matrix on random targets (MATK) + bit reversal ----
    k = int(rng.integers(1, min(n, 4) + 1))
    tq = [int(q) for q in rng.choice(n, size=k, replace=False)]
    mat = rng.normal(size=(2**k, 2**k)) + 1j * rng.normal(size=(2**k, 2**k))
    prog = Program(n, [Op(OpKind.MATK if k > 1 else OpKind.MAT1, tuple(tq), (), matrix=mat, label="rand")])
    got = engine.run(engine.CompiledProgram(prog), {}, psi, device=dev).cpu()
    wd = oracle.run(prog, {}, psi)
    if (got - wd).abs().max() > 1e-10 * max(1.0, wd.abs().max().item()):
        fail("dense", n, tq, (got - wd).abs().max().item())
    rev = engine.bit_reverse(psi.to(dev)).cpu()
    if not torch.equal(rev, oracle.invert_endianness_state(psi, n)):
        fail("bit_reverse", n)
    count["dense"] += 1
    # ---- sampling: bit-exact on supplied uniforms, incl. edges and sparse states ----
    st = psi.clone()
    if rng.random() < 0.4:  # sparse state: most probabilities exactly zero
        mask = torch.tensor(rng.random(size=st.shape) < 0.9)
        st[mask] = 0
        st[:, int(rng.integers(2**n))] = 1.0
        st = st / st.abs().pow(2).sum(1, keepdim=True).sqrt()
    S = int(rng.integers(1, 200))
    u = rng.random(size=(B, S))
    u[:, 0] = 0.0
    if S > 1:
        u[:, 1] = np.nextafter(1.0, 0.0)
    idx = engine.sample_indices(st.to(dev), torch.tensor(u)).cpu().numpy()
    wi = oracle.sample_indices(oracle.probabilities(st), u)
    if not np.array_equal(idx, wi):
        fail("sample_indices", n, int((idx != wi).sum()))
    count["sample"] += 1
print("ok:", count)
