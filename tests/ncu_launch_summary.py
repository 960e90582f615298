"""Development tool (not a test): summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python tests/ncu_launch_summary.py <launches.csv> > summary.md"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"^.*::", "", re.sub(r"\(.*$", "", r["Kernel Name"]))
    val = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = val / 1e3 if unit in ("ns", "nsecond") else val if unit in ("us", "usecond") else val * 1e3
    rows.append((name, us))
agg = defaultdict(list)
for n, us in rows:
    agg[n].append(us)
total = sum(us for _, us in rows)
print("| kernel | launches | total ms | share | min us | median us | max us |\n|---|---|---|---|---|---|---|")
for n, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    v.sort()
    print("| `%s` | %d | %.2f | %.1f %% | %.1f | %.1f | %.1f |" % (n, len(v), sum(v) / 1e3, 100 * sum(v) / total, v[0], v[len(v) // 2], v[-1]))
print("\n%d launches, %.1f ms of kernel time under ncu" % (len(rows), total / 1e3))
