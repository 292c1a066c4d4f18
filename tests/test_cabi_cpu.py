"""The C-ABI shared library loads and exports every symbol `include/b200sv.h` declares (no compute
calls: there is no GPU in the CPU container), its binary record layouts match the planner's numpy
dtypes, and without a device the library fails loudly instead of falling back to the CPU."""

from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest
import torch

import __graft_entry__ as ge
from qadence_b200 import cabi, plan

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib() -> C.CDLL:
    ge.build()
    return cabi.load()


def test_every_declared_symbol_is_exported(lib: C.CDLL) -> None:
    header = (ROOT / "include" / "b200sv.h").read_text()
    declared = set(re.findall(r"^(?:int|int64_t) (b200sv_\w+)\(", header, flags=re.M))
    assert len(declared) >= 16
    assert declared == set(cabi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.b200sv_abi_version() == 2


def test_record_layouts_match_header() -> None:
    header = (ROOT / "include" / "b200sv.h").read_text()
    for struct, dt in (("b200sv_gate", plan.GATE_DT), ("b200sv_round", plan.ROUND_DT), ("b200sv_sweep", plan.SWEEP_DT), ("b200sv_pterm", plan.PTERM_DT),
                       ("b200sv_tma", plan.TMA_DT)):
        m = re.search(r"typedef struct %s \{\s*/\* (\d+) bytes" % struct, header)
        assert m and int(m.group(1)) == dt.itemsize, struct
    assert C.sizeof(cabi.PlanStruct) == 88
    assert int(re.search(r"#define B200SV_LT_STRIDE (\d+)", header).group(1)) == plan.LT_STRIDE
    from qadence_b200 import embedding

    fields = re.search(r"typedef struct b200sv_enode \{\s*int32_t ([^;]+);", header).group(1).split(",")
    assert embedding.ENODE_DT.itemsize == 4 * len(fields) == 32


def test_fails_loudly_without_a_device(lib: C.CDLL) -> None:
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert lib.b200sv_device_sm_count() == -3  # B200SV_E_NODEVICE
    with pytest.raises(cabi.B200svError, match="no CPU fallback"):
        cabi.require_device()
    from qadence_b200 import engine
    from qadence_b200.program import hea_program

    with pytest.raises(cabi.B200svError):
        engine.run(engine.CompiledProgram(hea_program(3, 1)), {nm: torch.zeros(1) for nm in hea_program(3, 1).param_names})


def test_argument_validation_returns_negative_codes(lib: C.CDLL) -> None:
    assert lib.b200sv_init_zero_state(None, 1, 3, 1, None) == -1
    assert lib.b200sv_apply_dense(None, None, 1, 3, 1, None, None, 1, None) == -1
    assert lib.b200sv_sample(None, 1, 3, 1, None, None, None, 10, None, None) == -1
