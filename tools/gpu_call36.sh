#!/bin/bash
# round-2 GPU call 36: the committed state, once more: full GPU suite, smoke, default bench line
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu > $O/c36_gpu_suite.log 2>&1; echo "gpu suite rc=$?"; tail -2 $O/c36_gpu_suite.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > $O/c36_bench.json 2> $O/c36_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open("gpurun_out/c36_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"], d["e2e"]["ms_per_step"], d["parity"]["ok"], d["roofline"]["traffic"], d["roofline"]["traffic_source"][:60])
PY
