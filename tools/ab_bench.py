"""A/B the bench workload under different build-free switches (one gpurun call, a few seconds per variant).

    python tools/ab_bench.py                      # the default list below
    python tools/ab_bench.py B200SV_PLANNER=greedy "B200SV_TILE_ADJ_C128=11,3,2 B200SV_PERSISTENT=2"

Each argument is a space-separated list of NAME=VALUE environment settings for one variant; every variant runs
`bench.py --no-cpu-baseline --steps 3 --warmup 3` in its own process and one table line is printed:
ms/step (state-resident and end-to-end), sweeps, average adjoint / forward sweep time and bandwidth fraction.
Switches read by the library: B200SV_PLANNER (search | greedy), B200SV_TILE_{FWD,ADJ}_{C128,C64} = "m,r,low",
B200SV_PERSISTENT (0 | 1 double-buffered | 2 table reuse), B200SV_FORCE_GEN, B200SV_DBG (timing experiments: results invalid).
"""

from __future__ import annotations

import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
DEFAULT = ["", "B200SV_PLANNER=greedy", "B200SV_PERSISTENT=2", "B200SV_TILE_ADJ_C128=11,3,2", "B200SV_TILE_FWD_C128=12,4,2"]


def main() -> None:
    variants = sys.argv[1:] or DEFAULT
    print(f"{'variant':<48} {'ms/step':>8} {'e2e':>8} {'fwd':>4} {'adj':>4} {'adj ms':>8} {'adj %':>6} {'fwd ms':>8} {'fwd %':>6}")
    for v in variants:
        env = dict(os.environ)
        for kv in v.split():
            k, _, val = kv.partition("=")
            env[k] = val
        out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--no-cpu-baseline", "--steps", "3", "--warmup", "3"],
                             capture_output=True, text=True, env=env, cwd=ROOT)
        line = next((ln for ln in reversed(out.stdout.splitlines()) if ln.startswith("{")), None)
        if line is None:
            print(f"{v or '(default)':<48} FAILED: {out.stderr.strip().splitlines()[-1] if out.stderr.strip() else out.returncode}")
            continue
        d = json.loads(line)
        r = d["roofline"]
        print(f"{v or '(default)':<48} {d['ms_per_step']:8.1f} {d['e2e']['ms_per_step']:8.1f} {d['config']['fwd_sweeps']:4d} {d['config']['adj_sweeps']:4d} "
              f"{r['avg_launch_ms']:8.2f} {100 * r['frac']:6.1f} {r['forward_sweep']['avg_launch_ms']:8.2f} {100 * r['forward_sweep']['frac']:6.1f}")


if __name__ == "__main__":
    main()
