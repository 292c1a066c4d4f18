#!/bin/bash
# round-2 GPU call 13: new shared-memory layout (tiles first, no alignment slack -> more CTAs per SM), register targets,
# virtual permutations (B200SV_VPERM) — correctness first (lean_check, lean tests), then timings of each variant
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== lean_check product VPERM=1"; timeout 300 python tools/lean_check.py --n 14 --depth 3 --batch 3 2>&1 | tail -2
echo "== lean_check product VPERM=0"; B200SV_VPERM=0 timeout 300 python tools/lean_check.py --n 14 --depth 3 --batch 3 2>&1 | tail -2
echo "== lean_check product c64"; timeout 300 python tools/lean_check.py --n 14 --depth 3 --batch 3 --c64 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_lean.py tests/test_lindblad.py tests/test_gpu_parity.py -q -m gpu -x > $O/c13_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/c13_tests.log
for V in 1 0; do
  echo "== product (a128 f168) VPERM=$V"; B200SV_VPERM=$V python tools/profile_step.py --steps 3 2>&1 | grep -E "TIMES"
done
for L in lib_a96_f128 lib_a96_f112 lib_a128_f128_dbgx; do
  echo "== $L check"; B200SV_LIB=build_variants/$L.so timeout 300 python tools/lean_check.py --n 14 --depth 3 --batch 3 2>&1 | tail -1
  for V in 1 0; do
    echo "== $L VPERM=$V"; B200SV_LIB=build_variants/$L.so B200SV_VPERM=$V python tools/profile_step.py --steps 3 2>&1 | grep -E "TIMES"
  done
done
echo "== a96_f128 5 CTAs forced?"; B200SV_LIB=build_variants/lib_a96_f128.so B200SV_LEAN_CTAS_PER_SM=4 python tools/profile_step.py --steps 3 2>&1 | grep -E "TIMES"
echo "== c64 product"; python tools/profile_step.py --steps 3 --dtype c64 2>&1 | grep -E "TIMES"
echo "== c64 a96_f128"; B200SV_LIB=build_variants/lib_a96_f128.so python tools/profile_step.py --steps 3 --dtype c64 2>&1 | grep -E "TIMES"
