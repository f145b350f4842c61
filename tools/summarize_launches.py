#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, average and total
duration and share per kernel.  usage: summarize_launches.py launches.csv [header text] > summary.txt"""
import csv
import sys

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.reader(lines)
hdr = next(rd)
ix = {h: i for i, h in enumerate(hdr)}
agg = {}
for r in rd:
    if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
    val = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    us = val / 1e3 if unit.startswith("ns") else (val * 1e3 if unit.startswith("ms") else val)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(v[1] for v in agg.values())
if len(sys.argv) > 2:
    print(sys.argv[2])
for name, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-28s launches=%5d  avg %9.1f us  total %10.1f us  share %5.1f%%" % (name, cnt, us / cnt, us, 100 * us / tot))
print("%-28s launches=%5d  %s total %10.1f us" % ("ALL", sum(v[0] for v in agg.values()), " " * 16, tot))
