#!/bin/bash
# round-2 GPU call 25: tabulated diagonal for constant multi-Z observables; closing CNOT ladder absorbed (cfg2: 8 + 8 sweeps)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_clifford.py tests/test_zz_gpu_oracle_full.py tests/test_gpu_qadence_plugin.py tests/test_gpu_parity.py -q -m gpu > $O/c25_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/c25_tests.log
echo "== absorbed + table"; python tools/profile_step.py --steps 3 2>&1 | grep -E "TIMES|E\["
echo "== not absorbed"; B200SV_ABSORB_PERM=0 python tools/profile_step.py --steps 3 2>&1 | grep -E "TIMES|E\["
timeout 600 python bench.py --no-cpu-baseline --steps 5 > $O/c25_bench.json 2> $O/c25_bench.err; echo "bench rc=$?"; tail -2 $O/c25_bench.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/c25_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"], d["e2e"]["ms_per_step"], d["parity"], d["roofline"]["other_kernels_ms"], d["config"]["fwd_sweeps"], d["config"]["adj_sweeps"])
PY
