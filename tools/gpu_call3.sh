#!/bin/bash
# round-2 GPU call 3: TMA swizzle probe, per-tile timing traces (lockstep diagnosis), accumulator-split A/B
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
  env "$@" timeout 600 python bench.py --no-cpu-baseline --steps 3 > $O/c3_bench_$name.json 2> $O/c3_bench_$name.err
  echo "$name: $(summ $O/c3_bench_$name.json)" | tee -a $O/c3_bench_summary.txt
}
rm -f $O/c3_bench_summary.txt
timeout 120 ./tools/ubench/tma_probe > $O/c3_tma_probe.log 2>&1; cat $O/c3_tma_probe.log
B200SV_TMA_SWIZZLE=1 timeout 300 python tools/lean_check.py > $O/c3_check_swizzle.log 2>&1; tail -3 $O/c3_check_swizzle.log
B200SV_TMA_SWIZZLE=1 timeout 300 python tools/lean_check.py --c64 > $O/c3_check_swizzle_c64.log 2>&1; tail -3 $O/c3_check_swizzle_c64.log
run default X=1
run swizzle B200SV_TMA_SWIZZLE=1
B200SV_LEAN_TRACE=$O/c3_trace_s100 timeout 300 python tools/profile_step.py --steps 1 > $O/c3_trace1.log 2>&1
B200SV_LEAN_STAGGER=0 B200SV_LEAN_TRACE=$O/c3_trace_s0 timeout 300 python tools/profile_step.py --steps 1 > $O/c3_trace2.log 2>&1
python tools/lean_trace.py $O/c3_trace_s100.3.adj.bin $O/c3_trace_s100.2.fwd.bin $O/c3_trace_s0.3.adj.bin $O/c3_trace_s0.2.fwd.bin 2>&1 | tee $O/c3_trace_analysis.txt
rm -f $O/c3_trace_s100.[0145-7]*.bin $O/c3_trace_s0.[0145-7]*.bin
timeout 900 python -m pytest tests/test_gpu_qadence_plugin.py tests/test_gpu_peer.py tests/test_gpu_lean.py -q -m gpu > $O/c3_tests.log 2>&1; echo "tests rc=$?"; tail -6 $O/c3_tests.log
