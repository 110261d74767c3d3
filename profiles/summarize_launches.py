"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit, 1e-3)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"<.*", "", name)
        rows.append((name, v * scale))
tot = sum(v for _, v in rows)
agg = defaultdict(lambda: [0, 0.0])
for n, v in rows:
    agg[n][0] += 1
    agg[n][1] += v
print("launches=%d total=%.1f us" % (len(rows), tot))
print("%-48s %6s %12s %7s" % ("kernel", "count", "time_us", "share"))
for n, (c, v) in sorted(agg.items(), key=lambda t: -t[1][1]):
    print("%-48s %6d %12.1f %6.1f%%" % (n[:48], c, v, 100 * v / tot))
