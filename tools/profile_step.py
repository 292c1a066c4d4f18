"""One expectation+adjoint step of the bench workload (for ncu captures; prints nothing timed)."""
import ctypes as C
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from qadence_b200 import cabi, engine
from qadence_b200.engine import _DT, _stream_ptr
from qadence_b200.program import hea_program, total_magnetization

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
depth = int(sys.argv[3]) if len(sys.argv) > 3 else 8
dt = torch.complex128 if (len(sys.argv) <= 4 or sys.argv[4] == "c128") else torch.complex64
dev = torch.device("cuda:0")
prog = hea_program(n, depth)
cp = engine.CompiledProgram(prog)
cob = engine.CompiledObservable(total_magnetization(n))
vals = {nm: (torch.rand(B, dtype=torch.float64, device=dev) * 6.28).requires_grad_(True) for nm in cp.names}
for _ in range(2):
    out = engine.expectation(cp, [cob], vals, dtype=dt, device=dev)
    out.sum().backward()
torch.cuda.synchronize()
print("ok", out[:2, 0].tolist())
