#!/bin/bash
# round-2 GPU call 33: bench lines + launch list of the final planner setting (beam 6)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python bench.py > $O/c33_bench.json 2> $O/c33_bench.err; echo "bench rc=$?"; tail -c 200 $O/c33_bench.json
timeout 300 python tools/profile_step.py > $O/c33_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:sweep_kernel|pauli_|zero_state|resolve|gradient" -c 200 --csv --log-file $O/c33_launches.csv python tools/profile_step.py > $O/c33_ncu1.log 2>&1
timeout 600 python tools/e2e_breakdown.py 2>&1 | grep -E " ms " > $O/c33_e2e.log; cat $O/c33_e2e.log
