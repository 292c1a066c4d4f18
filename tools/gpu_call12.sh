#!/bin/bash
# round-2 GPU call 12: new tests (Lindblad, Trotter, dropout through the plugin) + timing experiments on the adjoint sweep:
# where the time of a round goes (B200SV_DBG bits of the experimental build) and r = 2 tiles (6 warps per scheduler)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > $O/c12_smi.txt
timeout 900 python -m pytest tests/test_lindblad.py tests/test_time_dependent.py tests/test_gpu_qadence_plugin.py tests/test_gpu_density.py -q -m gpu > $O/c12_new_tests.log 2>&1; echo "new tests rc=$?"; tail -4 $O/c12_new_tests.log
echo "== product"; python tools/profile_step.py --steps 3 2>&1 | grep TIMES
X=build_variants/libb200sv_x.so
echo "== experimental build, no flags"; B200SV_LIB=$X python tools/profile_step.py --steps 3 2>&1 | grep TIMES
for d in 1 4 5 8 12 13 16 32 20 28; do echo "== DBG=$d"; B200SV_LIB=$X B200SV_DBG=$d timeout 300 python tools/profile_step.py --steps 3 2>&1 | grep TIMES; done
echo "== r=2 m=10 adjoint"; B200SV_LIB=$X B200SV_TILE_ADJ_C128=10,2,2 timeout 300 python tools/profile_step.py --steps 3 2>&1 | grep -E "TIMES|E\[0"
echo "== r=2 m=9 adjoint"; B200SV_LIB=$X B200SV_TILE_ADJ_C128=9,2,2 timeout 300 python tools/profile_step.py --steps 3 2>&1 | grep -E "TIMES|E\[0"
echo "== r=3 m=10 reference values"; B200SV_LIB=$X timeout 300 python tools/profile_step.py --steps 3 2>&1 | grep -E "E\[0"
echo "== r=2 check vs generic"; B200SV_LIB=$X B200SV_TILE_ADJ_C128=10,2,2 timeout 300 python tools/lean_check.py --n 14 --depth 3 --batch 3 2>&1 | tail -3
