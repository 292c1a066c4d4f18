#!/bin/bash
# round-2 GPU call 29: backward pass without the state copy — tests, e2e breakdown, bench (e2e)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_qadence_plugin.py tests/test_second_order.py tests/test_gpu_run_grad.py tests/test_clifford.py -q -m gpu > $O/c29_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/c29_tests.log
timeout 600 python tools/e2e_breakdown.py 2>&1 | grep -E " ms " > $O/c29_e2e.log; cat $O/c29_e2e.log
timeout 600 python bench.py --no-cpu-baseline --steps 5 > $O/c29_bench.json 2> $O/c29_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open("gpurun_out/c29_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"], d["e2e"]["ms_per_step"], d["parity"]["ok"], d["roofline"]["traffic"])
PY
