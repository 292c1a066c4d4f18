"""Loader for tests/golden/cases/*.json (written by tests/golden/make_golden.py)."""

from __future__ import annotations

import json
from pathlib import Path

import torch

from qadence_b200.program import PauliSum, Program

CASES_DIR = Path(__file__).resolve().parent / "golden" / "cases"


def _cplx(d: dict) -> torch.Tensor:
    return (torch.tensor(d["re"], dtype=torch.float64) + 1j * torch.tensor(d["im"], dtype=torch.float64)).reshape(d["shape"])


def case_names() -> list[str]:
    return sorted(p.stem for p in CASES_DIR.glob("*.json"))


def load_case(name: str) -> dict:
    raw = json.loads((CASES_DIR / f"{name}.json").read_text())
    values = {k: torch.tensor(v, dtype=torch.float64) for k, v in raw["values"].items()}
    for k, v in raw.get("matrix_values", {}).items():
        values[k] = _cplx(v)
    case = {
        "name": raw["name"], "src": raw["src"], "n_qubits": raw["n_qubits"], "batch": raw["batch"],
        "program": Program.from_json(raw["program"]), "values": values,
        "state": _cplx(raw["state"]), "expected_state": _cplx(raw["expected_state"]),
    }
    if "observables" in raw:
        case["observables"] = [
            PauliSum.from_json(o) if kind == "PauliSum" else Program.from_json(o)
            for o, kind in zip(raw["observables"], raw["observable_kinds"])
        ]
        case["expected_expectation"] = torch.tensor(raw["expected_expectation"], dtype=torch.float64)
    if "expected_gate_grads" in raw:
        case["expected_gate_grads"] = {k: torch.tensor(v, dtype=torch.float64) for k, v in raw["expected_gate_grads"].items()}
    return case
