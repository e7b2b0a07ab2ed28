"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per kernel name, count and total time of the
LAST `1/n` of the launches (= the last of n identical steps).  python scripts/launch_summary.py file.csv n"""
import csv
import re
import sys
from collections import OrderedDict

path, n = sys.argv[1], int(sys.argv[2])
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1e-3)
        rows.append((r["Kernel Name"], v))
per = len(rows) // n
last = rows[len(rows) - per:]
agg = OrderedDict()
for name, us in last:
    name = re.sub(r"\(.*", "", name)
    c, t = agg.get(name, (0, 0.0))
    agg[name] = (c + 1, t + us)
tot = sum(t for _, t in agg.values())
print(f"{len(rows)} launches in file, {per} per step; last step: {tot / 1e3:.3f} ms of kernel time")
for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t / 1e3:8.3f} ms  {100 * t / tot:5.1f} %  x{c:<4d} {t / c:8.1f} us  {name[:100]}")
