"""cfg2 shape in complex64 (not the headline dtype): step time for the current tile defaults / env overrides."""
import sys, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from qadence_b200 import engine
from qadence_b200.program import hea_program, total_magnetization

n, depth, B = 20, 8, 256
dev = torch.device("cuda:0")
prog = hea_program(n, depth)
cp = engine.CompiledProgram(prog)
cob = engine.CompiledObservable(total_magnetization(n))
table = (torch.rand(len(cp.names), B, dtype=torch.float64, device=dev) * 6.28)
res = {}
for dt in (torch.complex64, torch.complex128):
    def step():
        t = table.clone().requires_grad_(True)
        e = engine.expectation(cp, [cob], {nm: t[i] for i, nm in enumerate(cp.names)}, dtype=dt, device=dev)
        e.sum().backward()
        return e
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        e = step()
    e1.record(); torch.cuda.synchronize()
    res[str(dt)] = {"ms_per_step": e0.elapsed_time(e1) / 3, "fwd_sweeps": cp.n_sweeps(dt, False), "adj_sweeps": cp.n_sweeps(dt, True), "E0": float(e[0, 0])}
print(json.dumps(res))
