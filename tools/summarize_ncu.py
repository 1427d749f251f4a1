"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
python tools/summarize_ncu.py gpurun_out/launches.csv [--last-step] > profiles/summary.txt
--last-step: only the launches of the last training step (the rows between the last two tavk::adamw_kernel launches:
the steady-state step when the driver ran several eager steps)."""
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
all_rows = [row for row in r if len(row) > vi]
if "--last-step" in sys.argv:
    ends = [i for i, row in enumerate(all_rows) if "adamw_kernel" in row[ki]]
    if len(ends) >= 2:
        all_rows = all_rows[ends[-2] + 1:ends[-1] + 1]
for row in all_rows:
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
fam = defaultdict(float)
for k, v in agg.items():
    key = ("gemm" if "gemm_bf16" in k else "attention" if "attn" in k else "layernorm" if "layernorm" in k else
           "adamw+sqnorm" if ("adamw" in k or "sqnorm" in k) else "other tavk" if "tavk::" in k else "non-tavk (ATen / memcpy / library)")
    fam[key] += v[1]
print("families: " + ", ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in sorted(fam.items(), key=lambda t: -t[1])))
lib = [k for k in agg if "cutlass" in k or "cudnn" in k or "cublas" in k.lower()]
print("library GEMM/conv kernels (cutlass*/cudnn*/cublas*): %s" % (lib if lib else "none"))
print("%-100s %7s %10s %7s" % ("kernel", "count", "total_us", "share"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print("%-100s %7d %10.0f %6.1f%%" % (k, v[0], v[1], 100 * v[1] / tot))
