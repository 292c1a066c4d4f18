#!/bin/bash
# round-2 GPU call 17: product after the geometry changes; experimental instances (forward r = 5, adjoint m = 9)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
X=build_variants/lib_exp.so
echo "== product"; python tools/profile_step.py --steps 3 2>&1 | grep -E "TIMES"
echo "== fwd 11,5"; B200SV_LIB=$X B200SV_TILE_FWD_C128=11,5,2 python tools/profile_step.py --steps 3 2>&1 | grep -E "TIMES|E\["
echo "== fwd 10,5"; B200SV_LIB=$X B200SV_TILE_FWD_C128=10,5,2 python tools/profile_step.py --steps 3 2>&1 | grep -E "TIMES|E\["
echo "== adj 9,4"; B200SV_LIB=$X B200SV_TILE_ADJ_C128=9,4,2 python tools/profile_step.py --steps 3 2>&1 | grep -E "TIMES|E\["
echo "== product E"; python tools/profile_step.py --steps 2 2>&1 | grep -E "E\["
echo "== product c64"; python tools/profile_step.py --steps 3 --dtype c64 2>&1 | grep -E "TIMES"
timeout 1200 python -m pytest tests/test_gpu_lean.py tests/test_gpu_parity.py tests/test_lindblad.py tests/test_gpu_full_size.py tests/test_gpu_edge_cases.py -q -m gpu > $O/c17_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/c17_tests.log
python bench.py --no-multi > $O/c17_bench.json 2> $O/c17_bench.err; echo "bench rc=$?"; cat $O/c17_bench.json
