"""The reference's OWN test files, unmodified, against this backend (replace mode + CPU oracle as the device).

A fast subset runs here; `python tools/run_reference_suite.py --report tests/golden/reference_suite_report.md` runs every
file of the reference suite that exercises the backend path (committed report: tests/golden/reference_suite_report.md).
"""

from __future__ import annotations

import sys
from pathlib import Path

import pytest

from tests.conftest import HAVE_REFERENCE

pytestmark = pytest.mark.reference
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tools"))


@pytest.mark.skipif(not HAVE_REFERENCE, reason="/root/reference not present")
@pytest.mark.parametrize("rel,min_passed", [("backends/test_endianness.py", 21), ("engines/test_torch.py", 8), ("qadence/test_matrices.py", 314),
                                            ("qadence/test_noise/test_digital_noise.py", 23)])
def test_reference_test_file_passes_on_b200sv(rel: str, min_passed: int) -> None:
    from run_reference_suite import run_file

    res = run_file(rel, timeout=600, extra=["--hypothesis-seed=0"])  # fixed examples: a reproducible gate for the CPU suite
    assert res["failed"] == 0 and res["errors"] == 0, res["failures"]
    assert res["passed"] >= min_passed, res
