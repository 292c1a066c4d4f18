#!/bin/bash
# round-2 GPU call 8: compute-sanitizer racecheck on the sweep kernels (small cases; ONE sanitizer tool per call)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
# plain runs first (must exit 0 without the tool)
timeout 300 python tools/lean_check.py --n 12 --depth 2 --batch 2 > $O/c8_plain_lean.log 2>&1; echo "plain lean rc=$?"
B200SV_TMA=0 timeout 300 python tools/lean_check.py --n 12 --depth 2 --batch 2 > $O/c8_plain_gather.log 2>&1; echo "plain gather rc=$?"
timeout 300 python tools/gpu_fuzz.py 5 > $O/c8_plain_fuzz.log 2>&1; echo "plain fuzz rc=$?"
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 20 python tools/lean_check.py --n 12 --depth 2 --batch 2 > $O/c8_racecheck_lean_tma.log 2>&1; echo "racecheck tma rc=$?"; tail -4 $O/c8_racecheck_lean_tma.log
B200SV_TMA=0 timeout 1200 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 20 python tools/lean_check.py --n 12 --depth 2 --batch 2 > $O/c8_racecheck_lean_gather.log 2>&1; echo "racecheck gather rc=$?"; tail -4 $O/c8_racecheck_lean_gather.log
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 20 python tools/gpu_fuzz.py 3 > $O/c8_racecheck_fuzz.log 2>&1; echo "racecheck fuzz rc=$?"; tail -4 $O/c8_racecheck_fuzz.log
