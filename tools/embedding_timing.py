"""Host-side cost of feeding 592 gate parameters (hea(16,12) + 16 feature-map gates, batch 512) to the engine:
(a) host dict of per-gate tensors -> pack_params (+ backward), (b) device-side embedding (+ backward).  Wall clock, GPU synced."""
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import sympy
import torch

from qadence_b200 import embedding as E
from qadence_b200 import engine
from qadence_b200.program import ProgramBuilder, hea_program

dev = torch.device("cuda:0")
n, B = 16, 512
b = ProgramBuilder(n)
for q in range(n):
    b.RX(q, f"fm{q}")
prog = b.build()
prog.ops += hea_program(n, 12).ops
cp = engine.CompiledProgram(prog)
tnames = [nm for nm in prog.param_names if not nm.startswith("fm")]
theta = {nm: torch.rand(1, dtype=torch.float64).requires_grad_(True) for nm in tnames}
x = torch.rand(B, dtype=torch.float64) * 1.98 - 0.99
xs = sympy.Symbol("x")
eprog = E.compile_expressions([(f"fm{q}", (q + 1) * sympy.acos(xs)) for q in range(n)] + [(nm, sympy.Symbol(nm)) for nm in tnames], lambda nm: nm == "x")
w = torch.rand(len(cp.names), B, dtype=torch.float64, device=dev)


def host():
    vals = {f"fm{q}": (q + 1) * torch.acos(x) for q in range(n)}
    vals.update(theta)
    tab = engine.pack_params(cp.names, vals, B, dev)
    (tab * w).sum().backward()


def device():
    var = torch.cat([theta[nm] for nm in eprog.var_names]).to(dev)
    tab = E.ParamTable(eprog.names, E.evaluate(eprog, var, x.reshape(1, B).to(dev), B), {}, torch.device("cpu"))
    tab = engine.pack_params(cp.names, tab, B, dev)
    (tab * w).sum().backward()


def wall(fn, reps=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


print(json.dumps({"slots": len(cp.names), "batch": B, "host_dict_pack_fwd_bwd_ms": wall(host), "device_embedding_fwd_bwd_ms": wall(device)}))
