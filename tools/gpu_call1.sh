#!/bin/bash
# round-2 GPU call 1: lean kernel correctness (diagnostic + tests), first bench numbers, A/B of geometries
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/c1_smi.txt 2>&1
echo "== lean_check tma nbuf1" > $O/c1_check.log
timeout 300 python tools/lean_check.py >> $O/c1_check.log 2>&1; echo "rc=$?" >> $O/c1_check.log
echo "== lean_check gather" >> $O/c1_check.log
B200SV_TMA=0 timeout 300 python tools/lean_check.py >> $O/c1_check.log 2>&1; echo "rc=$?" >> $O/c1_check.log
echo "== lean_check tma nbuf2" >> $O/c1_check.log
B200SV_LEAN_NBUF=2 timeout 300 python tools/lean_check.py >> $O/c1_check.log 2>&1; echo "rc=$?" >> $O/c1_check.log
echo "== lean_check c64" >> $O/c1_check.log
timeout 300 python tools/lean_check.py --c64 >> $O/c1_check.log 2>&1; echo "rc=$?" >> $O/c1_check.log
tail -5 $O/c1_check.log
timeout 900 python -m pytest tests/test_gpu_lean.py -q -m gpu -x > $O/c1_lean_tests.log 2>&1; echo "lean tests rc=$?"
tail -5 $O/c1_lean_tests.log
# bench: default geometry, then variations (short runs)
timeout 600 python bench.py --no-cpu-baseline --steps 3 > $O/c1_bench_default.json 2> $O/c1_bench_default.err; echo "bench rc=$?"
B200SV_LEAN_NBUF=2 timeout 600 python bench.py --no-cpu-baseline --steps 3 > $O/c1_bench_nbuf2.json 2> $O/c1_bench_nbuf2.err
B200SV_TILE_ADJ_C128=11,4,2 timeout 600 python bench.py --no-cpu-baseline --steps 3 > $O/c1_bench_adj_r4m11.json 2> $O/c1_bench_adj_r4m11.err
B200SV_TILE_ADJ_C128=11,3,2 timeout 600 python bench.py --no-cpu-baseline --steps 3 > $O/c1_bench_adj_r3m11.json 2> $O/c1_bench_adj_r3m11.err
B200SV_TILE_FWD_C128=12,5,2 timeout 600 python bench.py --no-cpu-baseline --steps 3 > $O/c1_bench_fwd_r5m12.json 2> $O/c1_bench_fwd_r5m12.err
B200SV_TILE_FWD_C128=10,4,2 timeout 600 python bench.py --no-cpu-baseline --steps 3 > $O/c1_bench_fwd_r4m10.json 2> $O/c1_bench_fwd_r4m10.err
B200SV_TMA=0 timeout 600 python bench.py --no-cpu-baseline --steps 3 > $O/c1_bench_gather.json 2> $O/c1_bench_gather.err
B200SV_LEAN=0 timeout 600 python bench.py --no-cpu-baseline --steps 3 > $O/c1_bench_oldkernel.json 2> $O/c1_bench_oldkernel.err
B200SV_DBG=1 timeout 600 python bench.py --no-cpu-baseline --steps 3 > $O/c1_bench_dbg1_nomath.json 2> $O/c1_bench_dbg1.err
B200SV_DBG=2 timeout 600 python bench.py --no-cpu-baseline --steps 3 > $O/c1_bench_dbg2_norounds.json 2> $O/c1_bench_dbg2.err
for f in $O/c1_bench_*.json; do echo "$f: $(python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d["roofline"]
    print(f"{d['ms_per_step']:.1f} ms/step e2e {d['e2e']['ms_per_step']:.1f}; adj {r['avg_launch_ms']:.2f} ms ({r['frac']:.3f}) x{d['config']['adj_sweeps']}; fwd {r['forward_sweep']['avg_launch_ms']:.2f} ms ({r['forward_sweep']['frac']:.3f}) x{d['config']['fwd_sweeps']}; fp64 {r['fp64_pipe']['frac']:.3f}; clk {d['clocks'].get('sm_mhz')}")
except Exception as e:
    print("unreadable", e)
PY
)"; done | tee $O/c1_bench_summary.txt
# the whole GPU suite
timeout 2400 python -m pytest tests -q -m gpu -x > $O/c1_gpu_suite.log 2>&1; echo "gpu suite rc=$?"
tail -5 $O/c1_gpu_suite.log
# ncu: launch list of two steps, then a full capture of one forward and one adjoint lean sweep
timeout 300 python tools/profile_step.py > $O/c1_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:sweep_kernel|pauli_|zero_state|resolve|gradient" -c 200 --csv --log-file $O/c1_launches.csv python tools/profile_step.py > $O/c1_ncu1.log 2>&1
timeout 300 python tools/profile_step.py --steps 1 > $O/c1_plain2.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:lean_sweep_kernel -s 2 -c 1 -o $O/c1_prof_fwd python tools/profile_step.py --steps 1 > $O/c1_ncu2.log 2>&1
timeout 300 python tools/profile_step.py --steps 1 > $O/c1_plain3.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:lean_sweep_kernel -s 10 -c 1 -o $O/c1_prof_adj python tools/profile_step.py --steps 1 > $O/c1_ncu3.log 2>&1
ls -la $O | grep c1_
