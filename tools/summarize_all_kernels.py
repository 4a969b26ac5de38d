#!/usr/bin/env python
"""gpurun_out/<tag>_all_{bench,tests}.csv (tools/capture_all_kernels.sh) -> profiles/<tag>_all_kernels.txt: one row per kernel:
launches, total and mean device time, warp instructions, DRAM bytes read / written, mean achieved occupancy and issue-active,
registers.  Times are under ncu (serialised, cold cache): compare shares and per-launch characteristics, not absolutes."""
import csv
import collections
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
rows = collections.defaultdict(lambda: collections.defaultdict(list))
SCALE = {"nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0,
         "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for part in ("bench", "tests"):
    p = os.path.join(ROOT, "gpurun_out", "%s_all_%s.csv" % (tag, part))
    if not os.path.exists(p):
        continue
    lines = [l for l in open(p, errors="replace") if l.startswith('"')]
    for r in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("<unnamed>::", "")
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        v *= SCALE.get(r["Metric Unit"], 1.0)
        rows[(part, name)][r["Metric Name"]].append(v)
out = ["# one row per kernel; from tools/capture_all_kernels.sh %s (light ncu pass; times are serialised, cold-cache)" % tag,
       "%-6s %-46s %7s %11s %11s %13s %11s %11s %6s %6s %5s" % ("where", "kernel", "launch", "total ms", "mean us", "warp inst", "dram rd MB", "dram wr MB", "occ%", "iss%", "regs")]
for (part, name), m in sorted(rows.items(), key=lambda kv: -sum(kv[1].get("gpu__time_duration.sum", [0]))):
    t = m.get("gpu__time_duration.sum", [0])
    mean = lambda k: sum(m.get(k, [0])) / max(1, len(m.get(k, [0])))
    out.append("%-6s %-46s %7d %11.3f %11.1f %13.3e %11.1f %11.1f %6.1f %6.1f %5d" % (
        part, name[:46], len(t), sum(t) * 1e3, mean("gpu__time_duration.sum") * 1e6, sum(m.get("smsp__inst_executed.sum", [0])),
        sum(m.get("dram__bytes_read.sum", [0])) / 1e6, sum(m.get("dram__bytes_write.sum", [0])) / 1e6,
        mean("sm__warps_active.avg.pct_of_peak_sustained_active"), mean("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        int(mean("launch__registers_per_thread"))))
dst = os.path.join(ROOT, "profiles", "%s_all_kernels.txt" % tag)
open(dst, "w").write("\n".join(out) + "\n")
print("\n".join(out[:40]))
