#!/bin/bash
# round-2 GPU call 15: adjoint r = 4 under dynamic scheduling (m = 11: 2 CTAs x 128 thr; m = 10: 4 CTAs x 64 thr), new c64 defaults
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
X=build_variants/lib_exp.so
echo "== product"; python tools/profile_step.py --steps 3 2>&1 | grep -E "TIMES"
echo "== adj 11,4 (product instance)"; B200SV_TILE_ADJ_C128=11,4,2 python tools/profile_step.py --steps 3 2>&1 | grep -E "TIMES"
echo "== adj 10,4 (experimental)"; B200SV_LIB=$X B200SV_TILE_ADJ_C128=10,4,2 python tools/profile_step.py --steps 3 2>&1 | grep -E "TIMES|E\["
echo "== adj 10,4 check"; B200SV_LIB=$X B200SV_TILE_ADJ_C128=10,4,2 timeout 300 python tools/lean_check.py --n 14 --depth 3 --batch 3 2>&1 | tail -1
echo "== adj 11,3"; B200SV_TILE_ADJ_C128=11,3,2 python tools/profile_step.py --steps 3 2>&1 | grep -E "TIMES"
echo "== c64 new defaults"; python tools/profile_step.py --steps 3 --dtype c64 2>&1 | grep -E "TIMES"
echo "== c64 check"; timeout 300 python tools/lean_check.py --n 14 --depth 3 --batch 3 --c64 2>&1 | tail -1
echo "== c64 adj 12,4 fwd 12,5"; B200SV_TILE_ADJ_C64=12,4,2 python tools/profile_step.py --steps 3 --dtype c64 2>&1 | grep -E "TIMES"
timeout 900 python -m pytest tests/test_gpu_lean.py tests/test_gpu_parity.py -q -m gpu -x > $O/c15_tests.log 2>&1; echo "tests rc=$?"; tail -2 $O/c15_tests.log
