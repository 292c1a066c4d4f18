"""The CPU oracle against the committed golden fixtures (reference front-end + reference dense
oracle `block_to_tensor`, see tests/golden/make_golden.py).  Runs anywhere — no qadence needed."""

from __future__ import annotations

import pytest
import torch

from oracle import statevec as oracle
from tests.golden_cases import case_names, load_case

NAMES = case_names()


def test_fixtures_present() -> None:
    assert len(NAMES) >= 16


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_golden(name: str) -> None:
    c = load_case(name)
    got = oracle.run(c["program"], c["values"], c["state"])
    assert (got - c["expected_state"]).abs().max().item() < 1e-10, c["src"]
    if "observables" in c:
        e = oracle.expectation(c["program"], c["observables"], c["values"], c["state"])
        assert (e - c["expected_expectation"]).abs().max().item() < 1e-10
    if "expected_gate_grads" in c:
        _, grads = oracle.adjoint_expectation_and_grads(c["program"], c["observables"][0], c["values"], c["state"])
        for k, v in c["expected_gate_grads"].items():
            assert (grads[k] - v).abs().max().item() < 1e-10
