"""CUDA path (through the C-ABI) against the golden fixtures generated from the reference's own
front-end and dense oracle.  complex128: 1e-10; complex64: 1e-5 (BASELINE.json tolerances)."""

from __future__ import annotations

import pytest
import torch

from qadence_b200 import engine
from qadence_b200.program import PauliSum
from tests.golden_cases import case_names, load_case

pytestmark = pytest.mark.gpu
TOL = {torch.complex128: 1e-10, torch.complex64: 1e-5}


@pytest.mark.parametrize("dtype", [torch.complex128, torch.complex64])
@pytest.mark.parametrize("name", case_names())
def test_cuda_matches_golden(name: str, dtype: torch.dtype) -> None:
    c = load_case(name)
    tol = TOL[dtype]
    cp = engine.CompiledProgram(c["program"])
    got = engine.run(cp, c["values"], c["state"], dtype=dtype, device="cuda:0").cpu().to(torch.complex128)
    scale = max(1.0, c["expected_state"].abs().max().item())
    assert (got - c["expected_state"]).abs().max().item() <= tol * scale, c["src"]
    if "observables" not in c:
        return
    cobs = [engine.CompiledObservable(o) for o in c["observables"]]
    vals = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in c["values"].items()}
    e = engine.expectation(cp, cobs, vals, c["state"], dtype=dtype, device="cuda:0")
    escale = max(1.0, c["expected_expectation"].abs().max().item())
    assert (e.detach().cpu().double() - c["expected_expectation"]).abs().max().item() <= 10 * tol * escale
    if "expected_gate_grads" in c and isinstance(c["observables"][0], PauliSum) and e.requires_grad:
        e[:, 0].sum().backward()
        for k, want in c["expected_gate_grads"].items():
            g = vals[k].grad
            assert g is not None, k
            got_g = g.cpu() if g.numel() == want.numel() else g.cpu().expand_as(want)
            ref = want if g.numel() == want.numel() else want.sum(0, keepdim=True).expand_as(want)
            gs = max(1.0, want.abs().max().item())
            assert (got_g - ref).abs().max().item() <= 100 * tol * gs, (k, c["src"])


def test_cfg1_known_shape() -> None:
    """BASELINE.json configs[0]: QNN hea(4,2) + total_magnetization, batch 16."""
    c = load_case("cfg1_qnn_hea4_depth2_batch16")
    cp = engine.CompiledProgram(c["program"])
    e = engine.expectation(cp, [engine.CompiledObservable(c["observables"][0])], c["values"], None, device="cuda:0")
    assert e.shape == (16, 1) and e.dtype == torch.float64
    assert cp.n_sweeps() == 1  # the whole 34-gate circuit is one kernel launch
