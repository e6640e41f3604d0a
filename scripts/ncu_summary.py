#!/usr/bin/env python
"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    if row["Metric Unit"] == "us":
        v *= 1e3
    elif row["Metric Unit"] == "ms":
        v *= 1e6
    agg.setdefault(row["Kernel Name"][:70], []).append(v)
tot = sum(sum(v) for k, v in agg.items() if "lcgs_b200" in k or not k.startswith("void at::"))
print("%-72s %5s %12s %12s %7s" % ("kernel", "n", "mean_ns", "sum_ns", "share"))
for k, v in agg.items():
    ours = not k.startswith("void at::")
    print("%-72s %5d %12.0f %12.0f %6.1f%%" % (k, len(v), sum(v) / len(v), sum(v), 100.0 * sum(v) / tot if ours else 0.0))
