#!/bin/bash
# round-2 GPU call 11: tabulated-diagonal Krylov matvec (cfg3), einsum-on-GPU informative baseline, HamEvo tests
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_time_dependent.py tests/test_gpu_full_size.py tests/test_gpu_edge_cases.py tests/test_gpu_golden.py -q -m gpu -k "hamevo or krylov or cfg3 or time or analog or evo or golden" > $O/c11_hamevo_tests.log 2>&1; echo "hamevo tests rc=$?"; tail -4 $O/c11_hamevo_tests.log
timeout 900 python tools/configs_check.py > $O/c11_configs.json 2> $O/c11_configs.err; echo "configs rc=$?"; python -c "
import json; d=json.load(open('$O/c11_configs.json')); print(json.dumps(d['cfg3']))"
timeout 600 python tools/einsum_on_gpu.py > $O/c11_einsum_gpu.json 2> $O/c11_einsum_gpu.err; echo "einsum rc=$?"; cat $O/c11_einsum_gpu.json; tail -2 $O/c11_einsum_gpu.err
timeout 600 python -m pytest tests/test_gpu_qadence_plugin.py -q -m gpu > $O/c11_plugin.log 2>&1; echo "plugin rc=$?"; tail -2 $O/c11_plugin.log
