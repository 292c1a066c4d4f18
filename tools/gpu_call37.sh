#!/bin/bash
# round-2 GPU call 37: tiles per claim of the dynamic tile scheduler (B200SV_LEAN_CHUNK, default 4) for the r = 4 kernels
cd "$(dirname "$0")/.."
for c in 4 1 2 8 16; do echo "== chunk $c"; B200SV_LEAN_CHUNK=$c python tools/profile_step.py --steps 3 2>&1 | grep -E "TIMES"; done
