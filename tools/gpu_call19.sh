#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python tools/e2e_breakdown.py > gpurun_out/c19_e2e.log 2>&1; echo rc=$?; head -40 gpurun_out/c19_e2e.log
