"""Summarise an .ncu-rep: key raw metrics per launch + instruction mix + hottest source lines."""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "smsp__cycles_active.avg"]
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    print("kernel:", r[hdr.index("Kernel Name")][:60])
    for w in want:
        if w in hdr:
            print(f"  {w} = {r[hdr.index(w)]} {rows[1][hdr.index(w)]}")
    st = sorted(((float(r[hdr.index(h)]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for h in stalls), reverse=True)
    print("  stalls/issue:", ", ".join(f"{n}={v:.2f}" for v, n in st[:8]))
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
agg, ops, cur = collections.OrderedDict(), collections.Counter(), None
for r in rows[3:]:
    if len(r) < 8:
        continue
    if r[0] != "":
        cur = (r[0], r[1].strip()[:95])
        agg.setdefault(cur, [0, 0])
    elif cur is not None:
        try:
            agg[cur][0] += int(r[6]); agg[cur][1] += int(r[7])
        except ValueError:
            continue
        op = r[3].strip().split()
        if op:
            o = op[1] if op[0].startswith("@") and len(op) > 1 else op[0]
            ops[o.split(".")[0]] += int(r[7])
ts = sum(v[0] for v in agg.values()) or 1
ti = sum(v[1] for v in agg.values()) or 1
print("instruction mix:", ", ".join(f"{o}={100*n/ti:.1f}%" for o, n in ops.most_common(14)))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[: int(sys.argv[2]) if len(sys.argv) > 2 else 16]:
    print(f"  L{k[0]:>4} samp {100*v[0]/ts:5.1f}% inst {100*v[1]/ti:5.1f}%  {k[1]}")
