#!/bin/bash
# round-2 GPU call 4: dynamic tile scheduling + batched moment reduction
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
run() { local name=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu-baseline --steps 3 > $O/c4_bench_$name.json 2> $O/c4_bench_$name.err
  echo "$name: $(summ $O/c4_bench_$name.json)" | tee -a $O/c4_bench_summary.txt
}
rm -f $O/c4_bench_summary.txt
timeout 300 python tools/lean_check.py > $O/c4_check.log 2>&1; tail -2 $O/c4_check.log
timeout 900 python -m pytest tests/test_gpu_lean.py tests/test_zz_gpu_oracle_full.py -q -m gpu -x > $O/c4_lean_tests.log 2>&1; echo "lean tests rc=$?"; tail -3 $O/c4_lean_tests.log
run default X=1
run chunk1 B200SV_LEAN_CHUNK=1
run chunk16 B200SV_LEAN_CHUNK=16
run fwd_r4m10 B200SV_TILE_FWD_C128=10,4,2
run adj_r4m11 B200SV_TILE_ADJ_C128=11,4,2
run gather B200SV_TMA=0
B200SV_LEAN_TRACE=$O/c4_trace timeout 300 python tools/profile_step.py --steps 1 > $O/c4_trace1.log 2>&1
python tools/lean_trace.py $O/c4_trace.3.adj.bin $O/c4_trace.2.fwd.bin 2>&1 | tee $O/c4_trace_analysis.txt
rm -f $O/c4_trace.[0145-7]*.bin
timeout 2400 python -m pytest tests -q -m gpu > $O/c4_gpu_suite.log 2>&1; echo "gpu suite rc=$?"; tail -5 $O/c4_gpu_suite.log
