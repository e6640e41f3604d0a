#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (one frame's worth).

usage: python scripts/launch_summary.py gpurun_out/launches_X.csv [frames]
"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = OrderedDict()
for r in rows[1:]:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = r[ix["Kernel Name"]].split("(")[0].replace("lcgs_b200::", "").replace("void ", "")
    v = float(r[ix["Metric Value"]].replace(",", ""))
    if r[ix["Metric Unit"]] in ("nsecond", "ns"):
        v /= 1000.0
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
for k, (n, t) in agg.items():
    print("%-70s launches %4d  total %9.1f us  avg %8.1f us  share %5.1f %%" % (k[:70], n, t, t / n, 100 * t / tot))
print("total %.1f us over all captured launches" % tot)
