#!/bin/bash
# round-2 GPU call 31: final state of the repository — full GPU suite, smoke, bench
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 2400 python -m pytest tests -q -m gpu > $O/c31_gpu_suite.log 2>&1; echo "gpu suite rc=$?"; tail -3 $O/c31_gpu_suite.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --no-cpu-baseline --steps 5 > $O/c31_bench.json 2> $O/c31_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open("gpurun_out/c31_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"], d["e2e"]["ms_per_step"], d["parity"]["ok"], d["roofline"]["traffic"], d["gpu_launches"])
PY
