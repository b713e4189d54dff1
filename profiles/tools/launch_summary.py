#!/usr/bin/env python
"""Average duration per kernel name from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, agg = None, {}
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        agg.setdefault(d["Kernel Name"][:60], []).append(float(d["Metric Value"].replace(",", "")))
for k, v in agg.items():
    print(f"{k:60s} x{len(v):3d}  avg {sum(v) / len(v) / 1e6:8.4f} ms")
