#!/bin/bash
# round-2 GPU call 26: final build (tabulated diagonal, absorbed ladder): full GPU suite, bench lines, ncu launch list + ONE full
# capture of the heaviest adjoint sweep (roofline.traffic)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
cat qadence_b200/lib/libb200sv.so.srchash > $O/c26_build_hash.txt
timeout 2400 python -m pytest tests -q -m gpu > $O/c26_gpu_suite.log 2>&1; echo "gpu suite rc=$?"; tail -4 $O/c26_gpu_suite.log
echo "== step"; python tools/profile_step.py --steps 3 2>&1 | grep -E "TIMES"
timeout 300 python tools/profile_step.py > $O/c26_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:sweep_kernel|pauli_|zero_state|resolve|gradient" -c 200 --csv --log-file $O/c26_launches.csv python tools/profile_step.py > $O/c26_ncu1.log 2>&1
timeout 300 python tools/profile_step.py --steps 1 > $O/c26_plain3.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:lean_sweep_kernel -s 14 -c 1 -o $O/c26_prof_adj python tools/profile_step.py --steps 1 > $O/c26_ncu3.log 2>&1
timeout 900 python bench.py > $O/c26_bench.json 2> $O/c26_bench.err; echo "bench rc=$?"; tail -c 300 $O/c26_bench.json; tail -2 $O/c26_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/c26_bench_ref.json 2> $O/c26_bench_ref.err; echo "ref rc=$?"
ls -la $O | grep c26_
