"""Hottest SASS instructions (by warp-stall samples) of one kernel of an ncu report, with the dominant stall reason.
python tools/ncu_src_hot.py report.ncu-rep <launch index (1-based)> [top N]"""
import csv
import io
import subprocess
import sys

rep, idx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
starts = [i for i, l in enumerate(lines) if l.startswith('"Kernel Name"')]
k0 = starts[int(idx) - 1]
k1 = starts[int(idx)] if int(idx) < len(starts) else len(lines)
print("launch %s of %d: %s" % (idx, len(starts), lines[k0][:170]))
rows = list(csv.reader(io.StringIO("\n".join(lines[k0 + 1:k1]))))
head = rows[0]
body = [r for r in rows[1:] if len(r) == len(head)]
si = head.index("# Samples")
src = head.index("Source")
ex = head.index("Instructions Executed")
stalls = [(i, h) for i, h in enumerate(head) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[si] or 0) for r in body)
print("total samples %d, instructions %d" % (tot, len(body)))
agg = {}
for i, h in stalls:
    agg[h] = sum(int(r[i] or 0) for r in body)
print("stall totals: " + ", ".join("%s=%.1f%%" % (h[6:], 100.0 * v / max(1, sum(agg.values()))) for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
order = sorted(range(len(body)), key=lambda k: -int(body[k][si] or 0))[:top]
for k in sorted(order):
    r = body[k]
    best = max(stalls, key=lambda ih: int(r[ih[0]] or 0))
    print("%5d %6.2f%% exec=%-8s %-22s %s" % (k, 100.0 * int(r[si] or 0) / tot, r[ex], best[1][6:], r[src].strip()[:110]))
