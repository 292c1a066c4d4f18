"""Run the REFERENCE's own, unmodified test files against the b200sv backend (CPU container; test infrastructure).

    python tools/run_reference_suite.py [--files a.py b.py ...] [--report tests/golden/reference_suite_report.md]

How: `tests/refsuite/refsuite_plugin.py` switches the plugin to replace mode (b200sv answers to the name "pyqtorch"), puts
the name-only shims of the absent third-party packages on the path and replaces the C-ABI-backed engine entry points by
the CPU oracle (there is no GPU here).  Tests parametrised on the pulser / horqrux backends are deselected.
"""

from __future__ import annotations

import argparse
import os
import re
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
REF_TESTS = Path("/root/reference/tests")

# the files of the reference suite that exercise the backend path, relative to /root/reference/tests
DEFAULT_FILES = [
    "backends/test_endianness.py",
    "backends/test_adjoint.py",
    "backends/test_backends.py",
    "backends/test_gpsr.py",
    "backends/test_utils.py",
    "backends/pyq/test_quantum_pyq.py",
    "backends/pyq/test_time_dependent_generator.py",
    "engines/test_torch.py",
    "qadence/test_matrices.py",
    "qadence/test_quantum_model.py",
    "qadence/test_overlap.py",
    "qadence/test_noise/test_digital_noise.py",
    "qadence/test_noise/test_readout.py",
    "analog/test_analog_emulation.py",
    "test_whitepaper.py",
    "test_finitediff.py",
    "constructors/test_ansatz.py",
    "constructors/test_daqc.py",
    "constructors/test_feature_maps.py",
    "constructors/test_hamiltonians.py",
    "constructors/test_qft.py",
    "constructors/test_rydberg_hea.py",
    "ml_tools/test_qnn.py",
]
# not in the default list: qadence/test_measurements/test_tomography.py and test_shadows.py (25 min of shot-based estimators on
# the CPU oracle; their backend contact is `Backend.sample`, covered by tests/test_qadence_adapter_cpu.py)
DESELECT = "not horqrux and not pulser and not HORQRUX and not PULSER and not jax and not JAX"


# failures that are not about this backend, with the reason (checked one by one)
NOT_APPLICABLE = [
    ("_Absent", "compares with / calls a package that is not in this image (horqrux, jax, qutip, pure pyqtorch)"),
    ("pulser", "pulser backend (not in this image)"),
    ("horqrux", "horqrux backend (not in this image)"),
    ("different_backends", "compares pyqtorch with horqrux (not in this image)"),
    ("NGOpt", "nevergrad is not in this image"),
    ("test_save_load_qm_pyq", "installed networkx serialises graphs with 'edges' instead of 'links': Register._from_dict of the reference fails before any backend is involved"),
    ("test_model_inputs_in_observable", "installed sympy 1.14 (the reference pins sympy<1.13.4) re-creates a Parameter on deepcopy (qadence/parameters.py:128 draws a new random value): `w != w` without any backend involved"),
    ("test_first_order_hea_derivatives[adjoint_observables1]", "same deepcopy effect: the un-valued observable parameter 'xobs' gets a new random value at every convert()"),
    ("test_constant_and_feature_transformed_module", "installed sympy 1.14 (the reference pins sympy<1.13.4) drops the `trainable=False` assumption when a Parameter is deep-copied: the non-trainable 'scale' / 'shift' inputs of the QNN come back as variational parameters before any backend is involved"),
    ("'torch.dtype' object has no attribute 'type'", "installed scipy / torch mismatch inside the reference's aGPSR shift optimiser"),
    ("FileNotFoundError", "test data path relative to the reference's repository root"),
]


def classify(failure: str) -> str:
    for needle, why in NOT_APPLICABLE:
        if needle in failure:
            return why
    return ""


def run_file(rel: str, timeout: int, extra: list[str], test_timeout: int = 300) -> dict:
    env = {**os.environ, "PYTHONPATH": f"{ROOT / 'tests' / 'refsuite'}:{ROOT}", "QADENCE_LOG_LEVEL": "ERROR"}
    cmd = [sys.executable, "-m", "pytest", "-p", "refsuite_plugin", "-p", "no:cacheprovider", rel, "-q", "--rootdir", str(REF_TESTS),
           "-W", "ignore", "-k", DESELECT, "-rfE", "--timeout", str(test_timeout), "--tb=line", *extra]
    t0 = time.time()
    try:
        out = subprocess.run(cmd, cwd=REF_TESTS, env=env, capture_output=True, text=True, timeout=timeout)
        text = out.stdout + out.stderr
    except subprocess.TimeoutExpired as e:
        text = (e.stdout or b"").decode() + "\nTIMEOUT"
    dt = time.time() - t0
    res = {"file": rel, "seconds": dt, "passed": 0, "failed": 0, "skipped": 0, "errors": 0, "deselected": 0, "failures": []}
    summary = next((ln for ln in reversed(text.splitlines()) if re.search(r"\d+ (passed|failed|error|errors|skipped|deselected)", ln) and " in " in ln), "")
    m = re.findall(r"(\d+) (passed|failed|skipped|deselected|error|errors|xfailed|xpassed)", summary)
    for num, what in m:
        key = {"error": "errors"}.get(what, what)
        res[key] = res.get(key, 0) + int(num)
    for line in text.splitlines():
        if line.startswith(("FAILED ", "ERROR ")):
            res["failures"].append(line[:300])
    if "TIMEOUT" in text[-20:]:
        res["failures"].append("TIMEOUT of the whole file")
    return res


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--files", nargs="*", default=DEFAULT_FILES)
    ap.add_argument("--report", default="")
    ap.add_argument("--timeout", type=int, default=1500, help="seconds per file")
    ap.add_argument("--test-timeout", type=int, default=300, help="seconds per test")
    ap.add_argument("pytest_args", nargs="*")
    args = ap.parse_args()
    rows = []
    for rel in args.files:
        if not (REF_TESTS / rel).exists():
            print(f"skip {rel}: not in the reference")
            continue
        r = run_file(rel, args.timeout, args.pytest_args, args.test_timeout)
        rows.append(r)
        print(f"{rel}: {r['passed']} passed, {r['failed']} failed, {r['errors']} errors, {r['skipped']} skipped, {r['deselected']} deselected ({r['seconds']:.0f} s)", flush=True)
        for f in r["failures"]:
            print("   ", f)
    for r in rows:
        real = [f for f in r["failures"] if f.startswith(("FAILED", "ERROR ")) and "::" in f]
        r["na"] = sum(1 for f in real if classify(f))
        r["ours"] = [f for f in real if not classify(f)]
    tot = {k: sum(r[k] for r in rows) for k in ("passed", "failed", "errors", "skipped", "deselected", "na")}
    tot["ours"] = sum(len(r["ours"]) for r in rows)
    print("TOTAL", tot)
    if args.report:
        lines = [
            "# The reference's own test files, unmodified, against b200sv",
            "",
            "Generated by `python tools/run_reference_suite.py --report tests/golden/reference_suite_report.md` in the CPU container:",
            "replace mode (`B200SV_REPLACE_PYQTORCH=1`: every `backend=\"pyqtorch\"` of the tests builds the b200sv backend), the CPU oracle",
            "stands in for the GPU kernels (tests/refsuite/refsuite_plugin.py). Tests parametrised on the pulser / horqrux backends are",
            "deselected. \"not applicable\" = fails for a reason that does not involve this backend (reason listed below).",
            "",
            "| reference test file | passed | failed: this backend | failed: not applicable | skipped | deselected (other backends) | s |",
            "|---|---:|---:|---:|---:|---:|---:|",
        ]
        for r in rows:
            lines.append(f"| `tests/{r['file']}` | {r['passed']} | {len(r['ours'])} | {r['na']} | {r['skipped']} | {r['deselected']} | {r['seconds']:.0f} |")
        lines.append(f"| **total** | **{tot['passed']}** | **{tot['ours']}** | {tot['na']} | {tot['skipped']} | {tot['deselected']} | |")
        lines += ["", "## Failures attributed to this backend", ""]
        ours = [f for r in rows for f in r["ours"]]
        lines += [f"* `{f}`" for f in ours] or ["none"]
        lines += ["", "## Not applicable", ""]
        seen: dict[str, list[str]] = {}
        for r in rows:
            for f in r["failures"]:
                why = classify(f)
                if why and "::" in f:
                    seen.setdefault(why, []).append(f.split(" - ")[0].replace("FAILED ", ""))
        for why, tests in seen.items():
            lines.append(f"* {why}: {len(tests)} — " + ", ".join(f"`{t}`" for t in tests[:6]) + (" …" if len(tests) > 6 else ""))
        Path(args.report).write_text("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
