"""Sums an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name (device time per launch, launches, share)."""
import csv
import re
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = v / 1e3 if unit.startswith("n") else (v if unit.startswith("u") else v * 1e3)
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    rows.append((name, us))
agg = OrderedDict()
for n, us in rows:
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
print("%-60s %6s %10s %10s %7s" % ("kernel", "n", "us/launch", "us total", "share"))
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-60s %6d %10.1f %10.1f %6.1f%%" % (n[:60], c, t / c, t, 100 * t / tot))
print("total %.1f us over %d launches" % (tot, len(rows)))
