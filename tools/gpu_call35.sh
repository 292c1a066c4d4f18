#!/bin/bash
# round-2 GPU call 35: ncu capture of the heaviest adjoint sweep of the final build (header comment changed the source hash),
# lean tests and a short bench of that build
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
cat qadence_b200/lib/libb200sv.so.srchash > $O/c35_build_hash.txt
timeout 300 python tools/profile_step.py --steps 1 > $O/c35_plain.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:lean_sweep_kernel -s 13 -c 1 -o $O/c35_prof_adj python tools/profile_step.py --steps 1 > $O/c35_ncu.log 2>&1
timeout 600 python -m pytest tests/test_gpu_lean.py tests/test_gpu_parity.py -q -m gpu > $O/c35_tests.log 2>&1; echo "tests rc=$?"; tail -2 $O/c35_tests.log
ls -la $O/c35_prof_adj.ncu-rep
