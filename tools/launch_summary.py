#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = None; data = []
for r in rows:
    if r and r[0] == "ID": hdr = r; continue
    if hdr and len(r) == len(hdr): data.append(r)
ix = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in data:
    name = r[ix["Kernel Name"]][:72]
    v = float(r[ix["Metric Value"]]); u = r[ix["Metric Unit"]]
    v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"| kernel | launches | total us | share | us/launch |\n|---|---|---|---|---|")
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"| {n} | {c/div:g} | {t/div:.1f} | {100*t/tot:.1f}% | {t/c:.1f} |")
print(f"total {tot/div:.1f} us")
