#!/bin/bash
# round-2 GPU call 20: ncu evidence of the r02d build: launch list of two steps, ONE full report (heaviest adjoint sweep),
# raw metrics of a heavy forward sweep as CSV; e2e breakdown after the slimmer input packing
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
cat qadence_b200/lib/libb200sv.so.srchash > $O/c20_build_hash.txt
timeout 600 python tools/e2e_breakdown.py 2>&1 | grep -E " ms " > $O/c20_e2e.log; cat $O/c20_e2e.log
timeout 300 python tools/profile_step.py > $O/c20_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:sweep_kernel|pauli_|zero_state|resolve|gradient" -c 200 --csv --log-file $O/c20_launches.csv python tools/profile_step.py > $O/c20_ncu1.log 2>&1
timeout 300 python tools/profile_step.py --steps 1 > $O/c20_plain3.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:lean_sweep_kernel -s 16 -c 1 -o $O/c20_prof_adj python tools/profile_step.py --steps 1 > $O/c20_ncu3.log 2>&1
timeout 1500 ncu --set full --clock-control none -k "regex:lean_sweep_kernel" -s 1 -c 1 --csv --page raw --log-file $O/c20_raw_fwd.csv python tools/profile_step.py --steps 1 > $O/c20_ncu2.log 2>&1
ls -la $O | grep c20_; du -sh $O
