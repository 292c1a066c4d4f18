"""Timing + sanity of the other BASELINE.json configs on one GPU (standalone builders; parity of the
same circuit shapes is covered by tests/golden at small sizes).

  cfg1  QNN hea(4,2) + feature map, batch 16            -> latency (evals/s)
  cfg3  12-atom Rydberg register: AnalogRX / AnalogInteraction / AnalogRY as HamEvo, batch 1024
  cfg4  feature map + hea(16,12), batch 512 per GPU     -> one training step (fwd + adjoint)
"""
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch

from qadence_b200 import engine
from qadence_b200.program import OpKind, PauliSum, PauliTerm, ProgramBuilder, hea_program, total_magnetization

dev = torch.device("cuda:0")
out = {}


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


# ---- cfg1 ----
n, B = 4, 16
b = ProgramBuilder(n)
for q in range(n):
    b.RX(q, "phi")
prog = b.build()
prog.ops += hea_program(n, 2).ops
cp = engine.CompiledProgram(prog)
cob = engine.CompiledObservable(total_magnetization(n))
vals = {nm: torch.rand(1, dtype=torch.float64, device=dev).requires_grad_(True) for nm in prog.param_names if nm != "phi"}
vals["phi"] = torch.linspace(-0.9, 0.9, B, dtype=torch.float64, device=dev)


def cfg1():
    e = engine.expectation(cp, [cob], vals, device=dev)
    e.sum().backward()


ms = timeit(cfg1, reps=50, warm=5)
out["cfg1"] = {"n": n, "batch": B, "ms_per_fwd_bwd": ms, "evals_per_s": B / (ms * 1e-3), "sweeps_fwd": cp.n_sweeps(), "note": "launch-latency bound (4 KiB state); one fused sweep per direction"}

# ---- cfg3: 3x4 lattice, spacing 8 um; H_int = sum C6/r^6 N_i N_j, drive Omega/2 (cos phi X - sin phi Y) ----
n, B = 12, 1024
coords = [(8.0 * (i // 4), 8.0 * (i % 4)) for i in range(n)]
C6 = 865723.02
terms_int = []
for i in range(n):
    for j in range(i + 1, n):
        r = np.hypot(coords[i][0] - coords[j][0], coords[i][1] - coords[j][1])
        c = C6 / r**6 / 4.0  # N_i N_j = (1 - Z_i - Z_j + Z_i Z_j) / 4
        terms_int += [PauliTerm(c, (), ()), PauliTerm(-c, (), ((i, "Z"),)), PauliTerm(-c, (), ((j, "Z"),)), PauliTerm(c, (), ((i, "Z"), (j, "Z")))]
h_int = PauliSum(n, terms_int).simplified()
omega = np.pi
h_rx = PauliSum(n, h_int.terms + [PauliTerm(omega / 2, (), ((q, "X"),)) for q in range(n)]).simplified()
h_ry = PauliSum(n, h_int.terms + [PauliTerm(omega / 2, (), ((q, "Y"),)) for q in range(n)]).simplified()
prog3 = ProgramBuilder(n).hamevo(h_rx, "ta").hamevo(h_int, "tt").hamevo(h_ry, "tb").build()
g = torch.Generator().manual_seed(0)
v3 = {"ta": torch.rand(B, generator=g, dtype=torch.float64) * np.pi / omega, "tt": torch.rand(B, generator=g, dtype=torch.float64),
      "tb": torch.rand(B, generator=g, dtype=torch.float64) * np.pi / omega}
cp3 = engine.CompiledProgram(prog3)
cob3 = engine.CompiledObservable(total_magnetization(n))
res = {}


def cfg3():
    res["e"] = engine.expectation(cp3, [cob3], v3, device=dev)


ms3 = timeit(cfg3, reps=3, warm=1)
# matvec count of one evaluation (the Krylov propagator's only state-sized kernel besides the fused Lanczos step)
_calls = {"pauli_apply": 0, "lanczos_step": 0}
_orig_apply = engine.pauli_apply


def _counting_apply(*a, **k):
    _calls["pauli_apply"] += 1
    return _orig_apply(*a, **k)


_orig_apply_diag = engine.pauli_apply_diag


def _counting_apply_diag(*a, **k):
    _calls["pauli_apply"] += 1
    return _orig_apply_diag(*a, **k)


engine.pauli_apply = _counting_apply
engine.pauli_apply_diag = _counting_apply_diag
cfg3()
torch.cuda.synchronize()
engine.pauli_apply, engine.pauli_apply_diag = _orig_apply, _orig_apply_diag
n_matvec = _calls["pauli_apply"]  # Krylov matvecs of the two non-diagonal HamEvo (the all-Z one is a single elementwise kernel)
S3 = B * (1 << n) * 16
# per Lanczos step: matvec 2S + lanczos_a 4S + lanczos_orth 3S + lanczos_next 3S algorithmic bytes
bytes_per_step = 12 * S3
psi = engine.run(cp3, v3, device=dev)
norm_err = float((engine.dot(psi, psi).real - 1).abs().max())
# parity of a batch slice against the dense-matrix oracle (n = 12 -> 4096 x 4096 expm is feasible for a few elements)
from oracle import statevec as oracle

sl = {k: v[:2] for k, v in v3.items()}
want = oracle.run(prog3, sl)
err = float((psi[:2].cpu() - want).abs().max())
out["cfg3"] = {"n": n, "batch": B, "ms": ms3, "hamevo_applications_per_s": 3 * B / (ms3 * 1e-3), "generator_terms": len(h_rx.terms),
               "matvecs_per_evaluation": n_matvec, "ms_per_lanczos_step": ms3 / max(1, n_matvec), "matvecs_per_s": n_matvec * B / (ms3 * 1e-3),
               "hbm_frac_of_the_lanczos_steps": (bytes_per_step * n_matvec / (ms3 * 1e-3) / 1e9) / 6542.1,
               "note": "state 64 MiB: a Lanczos step is four kernels over 12 S = 768 MiB of algorithmic traffic; the state and the Krylov basis of the last steps are partly L2-resident (126 MB)",
               "norm_err": norm_err, "max_abs_err_vs_dense_expm_oracle(first 2 batch elements)": err}

# ---- cfg4: feature map (RX(x) per qubit here) + hea(16, 12), 512 per GPU ----
n, B = 16, 512
b = ProgramBuilder(n)
for q in range(n):
    b.RX(q, "x")
prog4 = b.build()
prog4.ops += hea_program(n, 12).ops
cp4 = engine.CompiledProgram(prog4)
cob4 = engine.CompiledObservable(total_magnetization(n))
theta = {nm: torch.rand(1, dtype=torch.float64, device=dev).requires_grad_(True) for nm in prog4.param_names if nm != "x"}
x = torch.rand(B, dtype=torch.float64, device=dev) * 1.98 - 0.99
opt = torch.optim.Adam(theta.values(), lr=0.01)


def cfg4():
    opt.zero_grad()
    e = engine.expectation(cp4, [cob4], {**theta, "x": x}, device=dev)
    loss = ((e[:, 0] - x**2) ** 2).mean()
    loss.backward()
    opt.step()


ms4 = timeit(cfg4, reps=5, warm=2)
out["cfg4"] = {"n": n, "batch_per_gpu": B, "gates": len(prog4.ops), "ms_per_training_step": ms4, "samples_per_s_per_gpu": B / (ms4 * 1e-3),
               "sweeps_fwd": cp4.n_sweeps(), "sweeps_adj": cp4.n_sweeps(adjoint=True)}

# ---- cfg4 as the qadence plugin drives it: CPU nn.Parameters (one per gate), Chebyshev-tower feature map (q+1)*acos(x),
#      (a) host embedding: one small torch expression per gate parameter, dict -> pack -> H2D (what blocks/embedding.py does)
#      (b) device embedding: one torch.cat + one H2D + one kernel (qadence_b200/embedding.py)
import sympy

from qadence_b200 import embedding as E

b = ProgramBuilder(n)
for q in range(n):
    b.RX(q, f"fm{q}")
prog5 = b.build()
prog5.ops += hea_program(n, 12).ops
cp5 = engine.CompiledProgram(prog5)
tnames = [nm for nm in prog5.param_names if not nm.startswith("fm")]
theta_h = {nm: torch.rand(1, dtype=torch.float64).requires_grad_(True) for nm in tnames}
x_h = torch.rand(B, dtype=torch.float64) * 1.98 - 0.99
opt_h = torch.optim.Adam(theta_h.values(), lr=0.01)
xs = sympy.Symbol("x")
eprog = E.compile_expressions([(f"fm{q}", (q + 1) * sympy.acos(xs)) for q in range(n)] + [(nm, sympy.Symbol(nm)) for nm in tnames],
                              lambda nm: nm == "x")


def step_host():
    opt_h.zero_grad()
    vals = {f"fm{q}": (q + 1) * torch.acos(x_h) for q in range(n)}
    vals.update(theta_h)
    e = engine.expectation(cp5, [cob4], vals, device=dev)
    loss = ((e[:, 0].cpu() - x_h**2) ** 2).mean()
    loss.backward()
    opt_h.step()


def step_device():
    opt_h.zero_grad()
    var = torch.cat([theta_h[nm] for nm in eprog.var_names]).to(dev)
    feat = x_h.reshape(1, B).to(dev)
    pt = E.ParamTable(eprog.names, E.evaluate(eprog, var, feat, B), {}, torch.device("cpu"))
    e = engine.expectation(cp5, [cob4], pt, device=dev)
    loss = ((e[:, 0].cpu() - x_h**2) ** 2).mean()
    loss.backward()
    opt_h.step()


def wall(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


out["cfg4_plugin_like"] = {"what": "same circuit, CPU-resident per-gate nn.Parameters + Adam, wall-clock ms per training step",
                           "host_embedding_ms": wall(step_host), "device_embedding_ms": wall(step_device), "slots": eprog.n_slots}


# ---- cfg4, DQC variant: the loss contains dE/dx (create_graph=True), so the optimiser step needs second derivatives:
#      the adjoint gradient is differentiated by exact shifts of the 16 feature-map gates (2 x 16 more forward+adjoint passes)
def step_dqc():
    opt_h.zero_grad()
    var = torch.cat([theta_h[nm] for nm in eprog.var_names]).to(dev)
    feat = x_h.reshape(1, B).to(dev).requires_grad_(True)
    pt = E.ParamTable(eprog.names, E.evaluate(eprog, var, feat, B), {}, torch.device("cpu"))
    e = engine.expectation(cp5, [cob4], pt, device=dev)[:, 0]
    (dedx,) = torch.autograd.grad(e.sum(), feat, create_graph=True)
    loss = ((dedx[0] - feat[0].detach()) ** 2).mean() + (e[0] - 0.5) ** 2
    loss.backward()
    opt_h.step()


out["cfg4_dqc"] = {"what": "training step whose loss contains dE/dx (second-order differentiation), batch 512, wall-clock ms",
                   "ms_per_training_step": wall(step_dqc, reps=3, warm=1)}
print(json.dumps(out, indent=1))
