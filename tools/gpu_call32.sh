#!/bin/bash
# round-2 GPU call 32: planner beam width 6 (cfg2: 8 + 8 sweeps, 44 rounds, no absorption needed) — step times, tests, bench
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== beam 6"; python tools/profile_step.py --steps 3 2>&1 | grep -E "TIMES|E\["
echo "== c64"; python tools/profile_step.py --steps 3 --dtype c64 2>&1 | grep -E "TIMES"
timeout 1500 python -m pytest tests/test_gpu_lean.py tests/test_gpu_parity.py tests/test_zz_gpu_oracle_full.py tests/test_gpu_full_size.py tests/test_clifford.py tests/test_gpu_qadence_plugin.py -q -m gpu > $O/c32_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/c32_tests.log
timeout 600 python bench.py --no-cpu-baseline --steps 5 > $O/c32_bench.json 2> $O/c32_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open("gpurun_out/c32_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"], d["e2e"]["ms_per_step"], d["parity"]["ok"], d["roofline"]["traffic"], d["config"]["fwd_sweeps"], d["config"]["adj_sweeps"], d["roofline"]["step"]["rounds_fwd"], d["roofline"]["other_kernels_ms"])
PY
