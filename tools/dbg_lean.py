import torch, sys
sys.path.insert(0, "/root/repo")
from qadence_b200 import engine
from qadence_b200.program import hea_program, total_magnetization
from oracle import statevec as oracle
n, depth, batch = 6, 1, 2
prog = hea_program(n, depth); obs = total_magnetization(n)
g = torch.Generator().manual_seed(0)
values = {nm: (torch.rand(batch, generator=g, dtype=torch.float64)).requires_grad_(True) for nm in prog.param_names}
cp = engine.CompiledProgram(prog); cob = engine.CompiledObservable(obs)
out = engine.expectation(cp, [cob], values, device="cuda:0"); out.sum().backward()
e, gr = oracle.adjoint_expectation_and_grads(prog, obs, {k: v.detach() for k, v in values.items()})
k = prog.param_names
print("E err", (out[:,0].detach().cpu()-e).abs().max().item())
for nm in k[:4]+k[-3:]: print(nm, values[nm].grad.tolist(), gr[nm].tolist())
