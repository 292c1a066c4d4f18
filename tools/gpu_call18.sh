#!/bin/bash
# round-2 GPU call 18: checkpoint — full GPU suite + bench of the build with the new tile geometry, packed inputs in the plugin path
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 600 python bench.py --no-cpu-baseline --steps 5 > $O/c18_bench.json 2> $O/c18_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open("gpurun_out/c18_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"], d["e2e"]["ms_per_step"], d["parity"]["ok"], d["roofline"]["other_kernels_ms"])
PY
timeout 2400 python -m pytest tests -q -m gpu > $O/c18_gpu_suite.log 2>&1; echo "gpu suite rc=$?"; tail -5 $O/c18_gpu_suite.log
