import json, sys
for ln in sys.stdin:
    if ln.startswith("{"):
        d = json.loads(ln); r = d["roofline"]
        print(f'{d["ms_per_step"]:.1f} ms/step e2e {d["e2e"]["ms_per_step"]:.1f} | adj {r["avg_launch_ms"]:.2f} ms frac {r["frac"]:.3f} x{d["config"]["adj_sweeps"]} | fwd {r["forward_sweep"]["avg_launch_ms"]:.2f} ms frac {r["forward_sweep"]["frac"]:.3f} x{d["config"]["fwd_sweeps"]}' + (f' | light {r["light_sweep"]["avg_launch_ms"]:.2f} ms frac {r["light_sweep"]["frac"]:.3f}' if "light_sweep" in r else ""))
    elif "rror" in ln:
        print(ln.rstrip())
