#!/bin/bash
# round-2 GPU call 21: final-build bench lines (own arm + reference arm), plugin / embedding tests after the packed inputs
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python bench.py > $O/c21_bench.json 2> $O/c21_bench.err; echo "bench rc=$?"; tail -c 600 $O/c21_bench.json; tail -2 $O/c21_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/c21_bench_ref.json 2> $O/c21_bench_ref.err; echo "ref rc=$?"; tail -c 400 $O/c21_bench_ref.json
timeout 900 python -m pytest tests/test_gpu_qadence_plugin.py tests/test_embedding.py tests/test_gpu_run_grad.py tests/test_second_order.py -q -m gpu > $O/c21_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/c21_tests.log
timeout 600 python tools/configs_check.py > $O/c21_configs.json 2> $O/c21_configs.err; echo "configs rc=$?"; head -c 1500 $O/c21_configs.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
