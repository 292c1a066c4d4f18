#!/bin/bash
# round-2 GPU call 10: fused Lanczos (cfg3), configs check, HamEvo tests, final-build bench + adjoint capture for roofline.traffic
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_time_dependent.py tests/test_gpu_full_size.py tests/test_gpu_edge_cases.py -q -m gpu -k "hamevo or krylov or cfg3 or time or analog or evo" > $O/c10_hamevo_tests.log 2>&1; echo "hamevo tests rc=$?"; tail -4 $O/c10_hamevo_tests.log
timeout 900 python tools/configs_check.py > $O/c10_configs.json 2> $O/c10_configs.err; echo "configs rc=$?"; head -c 1800 $O/c10_configs.json
timeout 600 python bench.py --no-cpu-baseline --steps 5 > $O/c10_bench.json 2> $O/c10_bench.err; echo "bench rc=$?"; tail -c 400 $O/c10_bench.json
timeout 2400 python -m pytest tests -q -m gpu > $O/c10_gpu_suite.log 2>&1; echo "gpu suite rc=$?"; tail -5 $O/c10_gpu_suite.log
