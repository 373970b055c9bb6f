"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (cold-cache, serialised: compare SHARES).
    python tools/launch_summary.py gpurun_out/r02b_launches.csv > profiles/r02_launches_summary.txt"""
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1], newline="")) if len(r) > 5]
hdr = None
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr = r
        rows = rows[i + 1:]
        break
k, v, u = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = {}
for r in rows:
    name = re.sub(r"\(.*", "", r[k])
    name = re.sub(r"^void ", "", name)
    t = float(r[v].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[u], 1.0)
    a = tot.setdefault(name, [0.0, 0])
    a[0] += t; a[1] += 1
total = sum(a[0] for a in tot.values())
print("launch list: %s  (per-launch times are cold-cache and serialised under ncu; shares are what to compare with bench.py's kernel_time_share)" % sys.argv[1])
for name, (t, n) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print("%10.3f ms %5.1f%%  n=%4d  avg=%8.1f us  %s" % (t / 1e3, 100 * t / total, n, t / n, name[:110]))
print("total %.3f ms" % (total / 1e3))
