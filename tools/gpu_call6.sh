#!/bin/bash
# round-2 GPU call 6: the new bench.py (parity gate, plugin e2e, CPU legs) + ncu evidence (ONE full report: keep gpurun_out < 64 MiB)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/c6_smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_peer.py tests/test_gpu_lean.py -q -m gpu > $O/c6_tests.log 2>&1; echo "tests rc=$?"; tail -4 $O/c6_tests.log
timeout 900 python bench.py > $O/c6_bench.json 2> $O/c6_bench.err; echo "bench rc=$?"; tail -c 300 $O/c6_bench.json; tail -3 $O/c6_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/c6_bench_ref.json 2> $O/c6_bench_ref.err; echo "ref rc=$?"
timeout 300 python tools/profile_step.py > $O/c6_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:sweep_kernel|pauli_|zero_state|resolve|gradient" -c 200 --csv --log-file $O/c6_launches.csv python tools/profile_step.py > $O/c6_ncu1.log 2>&1
timeout 300 python tools/profile_step.py --steps 1 > $O/c6_plain3.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:lean_sweep_kernel -s 12 -c 1 -o $O/c6_prof_adj python tools/profile_step.py --steps 1 > $O/c6_ncu3.log 2>&1
timeout 300 python tools/profile_step.py --steps 1 > $O/c6_plain2.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none -k "regex:lean_sweep_kernel|pauli_|zero_state" -s 2 -c 1 --csv --page raw --log-file $O/c6_raw_fwd.csv python tools/profile_step.py --steps 1 > $O/c6_ncu2.log 2>&1
timeout 300 python tools/profile_step.py --steps 1 > $O/c6_plain4.log 2>&1 &&
timeout 900 ncu --set full --clock-control none -k "regex:pauli_|zero_state" -c 3 --csv --page raw --log-file $O/c6_raw_small.csv python tools/profile_step.py --steps 1 > $O/c6_ncu4.log 2>&1
ls -la $O | grep c6_; du -sh $O
