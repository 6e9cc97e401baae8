#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: count, mean and share per kernel."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
hdr, agg = None, collections.OrderedDict()
for r in rows:
    if r[0] == "ID":
        hdr = r
        continue
    if hdr is None:
        continue
    d = dict(zip(hdr, r))
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v, u = float(d["Metric Value"].replace(",", "")), d["Metric Unit"]
    v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
    a = agg.setdefault(d["Kernel Name"][:70], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values()) or 1.0
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{a[0]:5d} x {a[1] / a[0]:10.1f} us  {100 * a[1] / tot:5.1f} %  {k}")
