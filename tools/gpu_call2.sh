#!/bin/bash
# round-2 GPU call 2: stagger A/B, new tests (plugin on GPU, full-size oracle parity), whole suite
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d["roofline"]
    print(f"{d['ms_per_step']:.1f} ms/step e2e {d['e2e']['ms_per_step']:.1f}; adj {r['avg_launch_ms']:.2f} ms ({r['frac']:.3f}) x{d['config']['adj_sweeps']}; fwd {r['forward_sweep']['avg_launch_ms']:.2f} ms ({r['forward_sweep']['frac']:.3f}) x{d['config']['fwd_sweeps']}; fp64 {r['fp64_pipe']['frac']:.3f}; clk {d['clocks'].get('sm_mhz')}")
except Exception as e:
    print("unreadable", e)
PY
}
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu-baseline --steps 3 > $O/c2_bench_$name.json 2> $O/c2_bench_$name.err
  echo "$name: $(summ $O/c2_bench_$name.json)" | tee -a $O/c2_bench_summary.txt
}
rm -f $O/c2_bench_summary.txt
timeout 900 python -m pytest tests/test_gpu_lean.py -q -m gpu -x > $O/c2_lean_tests.log 2>&1; echo "lean tests rc=$?"; tail -3 $O/c2_lean_tests.log
run default X=1
run stagger0 B200SV_LEAN_STAGGER=0
run stagger50 B200SV_LEAN_STAGGER=50
run stagger200 B200SV_LEAN_STAGGER=200
run fwd_r4m10 B200SV_TILE_FWD_C128=10,4,2
run adj_r4m11 B200SV_TILE_ADJ_C128=11,4,2
run adj_r3m11 B200SV_TILE_ADJ_C128=11,3,2
run nbuf2 B200SV_LEAN_NBUF=2
run ctas2 B200SV_LEAN_CTAS_PER_SM=2
run ctas3 B200SV_LEAN_CTAS_PER_SM=3
timeout 900 python -m pytest tests/test_gpu_qadence_plugin.py -q -m gpu > $O/c2_plugin_tests.log 2>&1; echo "plugin tests rc=$?"; tail -15 $O/c2_plugin_tests.log
timeout 900 python -m pytest tests/test_zz_gpu_oracle_full.py -q -m gpu > $O/c2_oracle_full.log 2>&1; echo "oracle-full rc=$?"; tail -5 $O/c2_oracle_full.log
timeout 2400 python -m pytest tests -q -m gpu > $O/c2_gpu_suite.log 2>&1; echo "gpu suite rc=$?"; tail -12 $O/c2_gpu_suite.log
