#!/usr/bin/env python
"""Reads gpurun_out/{launches.csv, attn_full.ncu-rep} (scratch) and writes the compact, committed
summaries under profiles/:  <tag>_launches.txt, <tag>_attn_summary.txt, <tag>_attn_hot_sass.txt"""
import collections
import csv
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G, P = ROOT / "gpurun_out", ROOT / "profiles"
P.mkdir(exist_ok=True)

# ---- launch list
lines = [l for l in open(G / "launches.csv") if l.startswith('"')]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e6 if u == "ns" else v / 1e3 if u in ("us", "usecond") else v
    agg.setdefault(row["Kernel Name"].split("(")[0][-60:], []).append(v)
tot = sum(sum(v) for v in agg.values())
with open(P / f"{tag}_launches.txt", "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
    f.write("command: python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu   (our kernels only)\n")
    for k, v in agg.items():
        f.write(f"{k:62s} n={len(v):3d} mean={sum(v)/len(v):9.4f} ms  share={100*sum(v)/tot:6.2f}%\n")
print(open(P / f"{tag}_launches.txt").read())

# ---- full capture of the attention kernel
rep = G / "attn_full.ncu-rep"
raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keep = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg",
        "sm__cycles_elapsed.avg.per_second", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "lts__t_bytes.sum", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]
with open(P / f"{tag}_attn_summary.txt", "w") as f:
    f.write("ncu --set full --clock-control none --import-source on -k regex:carved_attn (bench.py HY-720p drop 0.7)\n")
    for i, h in enumerate(hdr):
        if h in keep:
            f.write(f"{h:86s} {units[i]:16s} {[r[i] for r in rows[2:]]}\n")
print(open(P / f"{tag}_attn_summary.txt").read())

src = subprocess.run(["ncu", "-i", str(rep), "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
h = rows[hi[0]]
body = rows[hi[0] + 1:(hi[1] - 1 if len(hi) > 1 else len(rows))]
ci = {n: i for i, n in enumerate(h)}
S, X, E = ci["# Samples"], ci["Source"], ci["Instructions Executed"]
tot = sum(int(r[S] or 0) for r in body)
byop = collections.Counter()
for r in body:
    t = r[X].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    byop[op] += int(r[S] or 0)
with open(P / f"{tag}_attn_hot_sass.txt", "w") as f:
    f.write(f"warp-state samples by opcode (total {tot}); then the 40 hottest SASS instructions\n")
    for k, v in byop.most_common(16):
        f.write(f"  {k:14s} {v:9d} {100*v/tot:5.1f}%\n")
    top = sorted(range(len(body)), key=lambda i: -int(body[i][S] or 0))[:40]
    for i in sorted(top):
        f.write(f"{i:5d} samples={body[i][S]:>8s} exec={body[i][E]:>11s}  {body[i][X][:90]}\n")
print(open(P / f"{tag}_attn_hot_sass.txt").read()[:3500])
