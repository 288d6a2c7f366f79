#!/usr/bin/env python
"""Sum an ncu source page (cuda,sass view) over named CUDA source line ranges of kernels.cuh.
usage: python profiles/ncu_ranges.py x.csv name:lo-hi [name:lo-hi ...]   (helper lines < 700 are reported as 'helpers')"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if r and r[0] == "Line No")
ci = {}
for i, n in enumerate(hdr):
    ci.setdefault(n, i)
ranges = []
for a in sys.argv[2:]:
    n, r = a.split(":")
    lo, hi = r.split("-")
    ranges.append((n, int(lo), int(hi)))
agg = {n: [0.0, 0.0] for n, _, _ in ranges}
agg["other"] = [0.0, 0.0]
cur = None
for r in rows[rows.index(hdr) + 1:]:
    if len(r) < len(hdr) or r[0] == "Line No":
        continue
    if r[0]:
        cur = int(r[0])
        continue
    if cur is None or r[2] in ("", "..."):
        continue
    try:
        inst = float(r[ci["Instructions Executed"]] or 0)
        smp = float(r[ci["# Samples"]] or 0)
    except ValueError:
        continue
    name = "other"
    for n, lo, hi in ranges:
        if lo <= cur <= hi:
            name = n
            break
    agg[name][0] += inst
    agg[name][1] += smp
ti = sum(v[0] for v in agg.values()) or 1
ts = sum(v[1] for v in agg.values()) or 1
for n, (i, s) in agg.items():
    print(f"{n:12s} {100 * i / ti:5.1f}% inst  {100 * s / ts:5.1f}% samples  ({i:.3e} warp-inst)")
