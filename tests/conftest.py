"""Test configuration.

* `gpu` marker: tests that need a real B200 (run by the driver with `-m gpu`).
* `reference` marker: tests that import the qadence front-end from /root/reference through the
  name-only shims under tests/_shims (CPU container only; skipped where the reference is absent,
  e.g. on the GPU box).
"""

from __future__ import annotations

import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
sys.dont_write_bytecode = True

# the qadence front-end: the read-only reference checkout in the CPU container, else the copy `__graft_entry__.build()`
# places under the git-ignored baseline/_ref/ (it travels to the GPU box with the snapshot)
REFERENCE = Path("/root/reference")
if not (REFERENCE / "qadence" / "__init__.py").exists():
    REFERENCE = ROOT / "baseline" / "_ref"
HAVE_REFERENCE = (REFERENCE / "qadence" / "__init__.py").exists()


def enable_reference() -> bool:
    """Make `import qadence` work (reference front-end + shims). Returns False when unavailable."""
    if not HAVE_REFERENCE:
        return False
    shims = str(ROOT / "tests" / "_shims")
    for p in (str(REFERENCE), shims):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.setdefault("QADENCE_LOG_LEVEL", "ERROR")
    return True


def pytest_configure(config: pytest.Config) -> None:
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    config.addinivalue_line("markers", "reference: needs /root/reference (qadence front-end) importable")


def pytest_collection_modifyitems(config: pytest.Config, items: list[pytest.Item]) -> None:
    import torch

    has_gpu = torch.cuda.is_available()
    skip_gpu = pytest.mark.skip(reason="no CUDA device")
    skip_ref = pytest.mark.skip(reason="qadence front-end not present (/root/reference or baseline/_ref)")
    for item in items:
        if "gpu" in item.keywords and not has_gpu:
            item.add_marker(skip_gpu)
        if "reference" in item.keywords and not HAVE_REFERENCE:
            item.add_marker(skip_ref)
