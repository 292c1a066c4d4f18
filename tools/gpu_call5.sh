#!/bin/bash
# round-2 GPU call 5: the new bench.py end to end (parity gate, plugin e2e, CPU legs) + ncu evidence of the final lean kernels
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/c5_smi.txt 2>&1
timeout 900 python bench.py > $O/c5_bench.json 2> $O/c5_bench.err; echo "bench rc=$?"; tail -c 600 $O/c5_bench.json; tail -3 $O/c5_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/c5_bench_ref.json 2> $O/c5_bench_ref.err; echo "ref rc=$?"
timeout 300 python tools/profile_step.py > $O/c5_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:sweep_kernel|pauli_|zero_state|resolve|gradient" -c 200 --csv --log-file $O/c5_launches.csv python tools/profile_step.py > $O/c5_ncu1.log 2>&1
timeout 300 python tools/profile_step.py --steps 1 > $O/c5_plain2.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:lean_sweep_kernel -s 2 -c 1 -o $O/c5_prof_fwd python tools/profile_step.py --steps 1 > $O/c5_ncu2.log 2>&1
timeout 300 python tools/profile_step.py --steps 1 > $O/c5_plain3.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:lean_sweep_kernel -s 11 -c 1 -o $O/c5_prof_adj python tools/profile_step.py --steps 1 > $O/c5_ncu3.log 2>&1
timeout 300 python tools/profile_step.py --steps 1 > $O/c5_plain4.log 2>&1 &&
timeout 900 ncu --set full --clock-control none -k "regex:pauli_" -c 2 -o $O/c5_prof_pauli python tools/profile_step.py --steps 1 > $O/c5_ncu4.log 2>&1
ls -la $O | grep c5_
