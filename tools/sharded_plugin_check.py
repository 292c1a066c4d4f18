"""`Configuration(shard="qubits")` end to end: expectation + gradients of a circuit on ONE register sharded over the process
group, through the qadence plugin when the front-end is importable (else through `qadence_b200.dist.sharded_expectation`),
against the single-GPU path; plus sampling on the sharded register.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 tools/sharded_plugin_check.py --qubits 18
"""
import argparse
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
import torch.distributed as dist

from qadence_b200 import dist as qdist
from qadence_b200 import engine
from qadence_b200.program import PauliSum, PauliTerm, hea_program

ap = argparse.ArgumentParser()
ap.add_argument("--qubits", type=int, default=18)
ap.add_argument("--depth", type=int, default=2)
ap.add_argument("--shots", type=int, default=4096)
args = ap.parse_args()
rank, world = qdist.init_from_env("nccl")
if not dist.is_initialized():  # world == 1 started without torchrun
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29577")
    dist.init_process_group("nccl", rank=0, world_size=1)
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
n = args.qubits
res = {"world": world, "n": n}

frontend = None
for cand in (Path("/root/reference"), ROOT / "baseline" / "_ref"):
    if (cand / "qadence" / "__init__.py").exists():
        sys.path[:0] = [str(cand), str(ROOT / "tests" / "_shims")]
        os.environ.setdefault("QADENCE_LOG_LEVEL", "ERROR")
        frontend = str(cand)
        break

obs_ps = PauliSum(n, [PauliTerm(1.0, (), ((q, "Z"),)) for q in range(n)] + [PauliTerm(0.7, (), ((0, "X"), (n - 1, "Y")))])
if frontend is not None:
    from qadence import QuantumCircuit, QuantumModel, X, Y, Z, add, hea, kron

    torch.manual_seed(0)
    circ = QuantumCircuit(n, hea(n, args.depth))
    qobs = add(*[Z(q) for q in range(n)], 0.7 * kron(X(0), Y(n - 1)))
    sharded = QuantumModel(circ, qobs, backend="b200sv", diff_mode="adjoint", configuration={"shard": "qubits"})
    single = QuantumModel(circ, qobs, backend="b200sv", diff_mode="adjoint")
    single.load_state_dict(sharded.state_dict())
    e_sh = sharded.expectation({})
    e_sh.sum().backward()
    g_sh = torch.stack([p.grad.reshape(()) for p in sharded.parameters() if p.requires_grad])
    e_1 = single.expectation({})
    e_1.sum().backward()
    g_1 = torch.stack([p.grad.reshape(()) for p in single.parameters() if p.requires_grad])
    res.update({"path": f"QuantumModel(backend='b200sv', configuration={{'shard': 'qubits'}}) (front-end {frontend})",
                "abs_err_expectation": float((e_sh - e_1).abs().max()), "max_abs_err_gradient": float((g_sh - g_1).abs().max()), "n_params": int(g_1.numel())})
    prog = sharded._circuit.native.program
    values = {nm: v for nm, v in sharded.embedding_fn(sharded._params, {}).items()}
else:
    prog = hea_program(n, args.depth)
    g = torch.Generator().manual_seed(1)
    values = {nm: (6.28 * torch.rand(1, generator=g, dtype=torch.float64)).requires_grad_(True) for nm in prog.param_names}
    e_sh = qdist.sharded_expectation(prog, [obs_ps], values, device=dev)
    e_sh.sum().backward()
    g_sh = torch.stack([values[nm].grad.reshape(()) for nm in prog.param_names])
    ref = {k: v.detach().clone().requires_grad_(True) for k, v in values.items()}
    e_1 = engine.expectation(engine.CompiledProgram(prog), [engine.CompiledObservable(obs_ps)], ref, device=dev)
    e_1.sum().backward()
    g_1 = torch.stack([ref[nm].grad.reshape(()) for nm in prog.param_names])
    res.update({"path": "qadence_b200.dist.sharded_expectation (qadence front-end absent)", "abs_err_expectation": float((e_sh.cpu() - e_1.cpu()).abs().max()),
                "max_abs_err_gradient": float((g_sh.cpu() - g_1.cpu()).abs().max()), "n_params": int(g_1.numel())})

# sampling on the sharded register vs the single-GPU sampler on the same uniforms
detached = {k: (v.detach() if isinstance(v, torch.Tensor) else v) for k, v in values.items()}
u = torch.rand(args.shots, dtype=torch.float64, generator=torch.Generator().manual_seed(3))
idx_sh = qdist._simulator(prog, torch.complex128, dev).sample(detached, args.shots, uniforms=u).cpu()
psi = engine.run(engine.CompiledProgram(prog), detached, device=dev)
idx_1 = engine.sample_indices(psi, u.reshape(1, -1)).reshape(-1).cpu()
res["sampling"] = {"shots": args.shots, "identical_indices": int((idx_sh == idx_1).sum()),
                   "note": "same uniforms; a shot can differ only when its uniform lies within rounding distance of a boundary of the cumulative distribution"}
if rank == 0:
    print(json.dumps(res))
for sim in list(qdist._SIMULATORS.values()):
    sim.close()
dist.barrier()
dist.destroy_process_group()
