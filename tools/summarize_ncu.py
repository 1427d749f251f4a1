"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
python tools/summarize_ncu.py gpurun_out/launches.csv > profiles/summary.txt"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
r = csv.reader(lines)
header = next(r)
ki, vi, ui = header.index("Kernel Name"), header.index("Metric Value"), header.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
n = 0
for row in r:
    if len(row) <= vi:
        continue
    val = float(row[vi].replace(",", ""))
    unit = row[ui]
    us = val / 1e3 if unit in ("ns", "nsecond") else (val if unit in ("us", "usecond") else val * 1e3)
    name = re.sub(r"\(.*", "", row[ki])[:100]
    agg[name][0] += 1
    agg[name][1] += us
    n += 1
tot = sum(v[1] for v in agg.values())
print("launches %d  total kernel time %.2f ms (cold-cache, serialised under ncu: compare SHARES)" % (n, tot / 1e3))
ours = sum(v[1] for k, v in agg.items() if "tavk::" in k)
print("tavk:: kernels: %.1f%% of kernel time, %d launches" % (100 * ours / tot, sum(v[0] for k, v in agg.items() if "tavk::" in k)))
print("%-100s %7s %10s %7s" % ("kernel", "count", "total_us", "share"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print("%-100s %7d %10.0f %6.1f%%" % (k, v[0], v[1], 100 * v[1] / tot))
