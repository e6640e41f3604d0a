#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output: stall-reason totals and the hottest SASS lines.

usage: ncu -i X.ncu-rep --page source --csv --kernel-name K --launch-count 1 > src.csv; ncu_stalls.py src.csv [N]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
data = []
for r in rows[hi + 1:]:
    if not r or not r[0].startswith("0x"):
        if data:
            break  # first kernel section only
        continue
    if len(r) == len(hdr):
        data.append(r)
ix = {h: i for i, h in enumerate(hdr)}
num = lambda r, k: int(float(r[ix[k]] or 0))  # noqa: E731
tot = sum(num(r, "# Samples") for r in data)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(num(r, s) for r in data) for s in stalls}
print("total samples", tot, "instructions", sum(num(r, "Instructions Executed") for r in data))
print({k: v for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v})
for r in sorted(data, key=lambda r: -num(r, "# Samples"))[:topn]:
    st = {s: num(r, s) for s in stalls if num(r, s)}
    st = dict(sorted(st.items(), key=lambda x: -x[1])[:3])
    print("%6d %-72s exec=%8d %s" % (num(r, "# Samples"), r[ix["Source"]].strip()[:72], num(r, "Instructions Executed"), st))
