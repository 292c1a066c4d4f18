#!/bin/bash
# round-2 GPU call 16: full-round fast path (product) vs without; forward at 144 registers; adjoint r = 4 default; tests
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== lean_check product"; timeout 300 python tools/lean_check.py --n 14 --depth 3 --batch 3 2>&1 | tail -1
echo "== lean_check product c64"; timeout 300 python tools/lean_check.py --n 14 --depth 3 --batch 3 --c64 2>&1 | tail -1
echo "== product (fast path, fwd 128, adj 10,4)"; python tools/profile_step.py --steps 3 2>&1 | grep -E "TIMES"
echo "== nofast"; B200SV_LIB=build_variants/lib_nofast.so python tools/profile_step.py --steps 3 2>&1 | grep -E "TIMES"
echo "== f144"; B200SV_LIB=build_variants/lib_f144.so python tools/profile_step.py --steps 3 2>&1 | grep -E "TIMES"
echo "== product adj 10,3"; B200SV_TILE_ADJ_C128=10,3,2 python tools/profile_step.py --steps 3 2>&1 | grep -E "TIMES"
echo "== nofast adj 10,3"; B200SV_LIB=build_variants/lib_nofast.so B200SV_TILE_ADJ_C128=10,3,2 python tools/profile_step.py --steps 3 2>&1 | grep -E "TIMES"
echo "== product c64"; python tools/profile_step.py --steps 3 --dtype c64 2>&1 | grep -E "TIMES"
echo "== nofast c64"; B200SV_LIB=build_variants/lib_nofast.so python tools/profile_step.py --steps 3 --dtype c64 2>&1 | grep -E "TIMES"
timeout 1200 python -m pytest tests/test_gpu_lean.py tests/test_gpu_parity.py tests/test_lindblad.py tests/test_gpu_adjoint.py -q -m gpu > $O/c16_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/c16_tests.log
