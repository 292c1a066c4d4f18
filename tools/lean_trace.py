#!/usr/bin/env python
"""Analyse the per-tile time stamps lean_sweep_kernel writes with B200SV_LEAN_TRACE=<prefix> (see launch_lean in
csrc/b200sv.cu): for every CTA the time a tile spends waiting for its load, in the rounds, and in the store / next-load
issue, and how the phases of the CTAs that share an SM line up.

    B200SV_LEAN_TRACE=gpurun_out/trace python tools/profile_step.py --steps 1 && python tools/lean_trace.py gpurun_out/trace.3.adj.bin
"""
import sys

import numpy as np


def main(path: str) -> None:
    raw = np.fromfile(path, dtype=np.uint64)
    grid, tiles, per_cta, nt = (int(x) for x in raw[:4].astype(np.int64))
    d = raw[4:].reshape(grid, tiles, 4).astype(np.int64)
    landed, done, issued, smid = d[..., 0], d[..., 1], d[..., 2], d[..., 3]
    ok = (landed > 0) & (done > 0) & (issued > 0)
    ntile = min(tiles, per_cta) - 1
    t0 = landed[:, 0].min()
    rounds = (done - landed)[:, 1:ntile]
    store = (issued - done)[:, 1:ntile]
    wait = (landed[:, 2:ntile + 1] - issued[:, 1:ntile])  # next tile landed - this tile's store/next-load issued
    period = landed[:, 2:ntile + 1] - landed[:, 1:ntile]
    print(f"{path}: grid {grid} ({nt} threads), {per_cta} tiles per CTA, traced {ntile}")
    print(f"  per tile [us]: period {period.mean()/1e3:.2f}  = rounds {rounds.mean()/1e3:.2f} + store/issue {store.mean()/1e3:.2f} + load wait {wait.mean()/1e3:.2f}")
    print(f"  percentiles of load wait [us]: {np.percentile(wait, [5, 50, 95]) / 1e3}")
    print(f"  percentiles of rounds    [us]: {np.percentile(rounds, [5, 50, 95]) / 1e3}")
    # phase alignment: for the CTAs of one SM, when (mod the period) does each tile land?
    sm_of = smid[:, 0]
    sms = np.unique(sm_of)
    print(f"  {len(sms)} SMs, CTAs per SM: {np.bincount(sm_of).max()}; slot (blockIdx // #SMs) by SM for SM {sms[0]}: {np.where(sm_of == sms[0])[0].tolist()}")
    spreads = []
    for s in sms[:148]:
        ctas = np.where(sm_of == s)[0]
        if len(ctas) < 2:
            continue
        k = min(20, ntile)
        ph = landed[ctas, k] - t0
        p = period[ctas].mean()
        rel = np.sort((ph - ph.min()) / p)
        spreads.append(rel)
    if spreads:
        sp = np.array([np.pad(r, (0, max(0, 8 - len(r))), constant_values=np.nan)[:8] for r in spreads])
        print("  landing times of tile 20 of the CTAs sharing an SM, relative to the first, in periods (mean over SMs):", np.nanmean(sp, axis=0).round(3))
    # global burstiness: how many CTAs are waiting for a load at the same instant
    ev = []
    for c in range(grid):
        for k in range(1, ntile):
            ev.append((issued[c, k], 1))
            ev.append((landed[c, k + 1], -1))
    ev.sort()
    cur, hist, last = 0, {}, ev[0][0]
    for t, dlt in ev:
        hist[cur] = hist.get(cur, 0) + (t - last)
        cur += dlt
        last = t
    tot = sum(hist.values())
    ks = sorted(hist)
    mean_wait = sum(k * v for k, v in hist.items()) / tot
    print(f"  CTAs waiting for a tile at the same time: mean {mean_wait:.1f} of {grid}; >50% of CTAs waiting during {100 * sum(v for k, v in hist.items() if k > grid / 2) / tot:.1f}% of the time")


if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
