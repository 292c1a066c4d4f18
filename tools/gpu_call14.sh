#!/bin/bash
# round-2 GPU call 14: forward tile geometry under the packed shared-memory layout; complex64 tile geometries
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== lean_check product"; timeout 300 python tools/lean_check.py --n 14 --depth 3 --batch 3 2>&1 | tail -1
echo "== product m=10 fwd (128 regs, 8 CTAs)"; python tools/profile_step.py --steps 3 2>&1 | grep -E "TIMES"
echo "== product m=11 fwd (128 regs, 4 CTAs x 128 thr)"; B200SV_TILE_FWD_C128=11,4,2 python tools/profile_step.py --steps 3 2>&1 | grep -E "TIMES"
echo "== f96 m=10 (10 CTAs)"; B200SV_LIB=build_variants/lib_f96.so python tools/profile_step.py --steps 3 2>&1 | grep -E "TIMES"
echo "== f96 m=11 (5 CTAs x 128 thr)"; B200SV_LIB=build_variants/lib_f96.so B200SV_TILE_FWD_C128=11,4,2 python tools/profile_step.py --steps 3 2>&1 | grep -E "TIMES"
echo "== f96 m=11 check"; B200SV_LIB=build_variants/lib_f96.so B200SV_TILE_FWD_C128=11,4,2 timeout 300 python tools/lean_check.py --n 14 --depth 3 --batch 3 2>&1 | tail -1
echo "== c64 default (fwd 12,4 adj 11,3)"; python tools/profile_step.py --steps 3 --dtype c64 2>&1 | grep -E "TIMES"
echo "== c64 adj 12,4"; B200SV_TILE_ADJ_C64=12,4,2 python tools/profile_step.py --steps 3 --dtype c64 2>&1 | grep -E "TIMES"
echo "== c64 adj 11,4"; B200SV_TILE_ADJ_C64=11,4,2 python tools/profile_step.py --steps 3 --dtype c64 2>&1 | grep -E "TIMES"
echo "== c64 fwd 12,5"; B200SV_TILE_FWD_C64=12,5,2 python tools/profile_step.py --steps 3 --dtype c64 2>&1 | grep -E "TIMES"
echo "== c64 fwd 11,4"; B200SV_TILE_FWD_C64=11,4,2 python tools/profile_step.py --steps 3 --dtype c64 2>&1 | grep -E "TIMES"
echo "== c64 DBG=1 (no math)"; B200SV_DBG=1 python tools/profile_step.py --steps 3 --dtype c64 2>&1 | grep -E "TIMES"
echo "== c64 DBG=2 (no rounds)"; B200SV_DBG=2 python tools/profile_step.py --steps 3 --dtype c64 2>&1 | grep -E "TIMES"
