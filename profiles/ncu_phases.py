#!/usr/bin/env python
"""Per-phase instruction totals of one kernel from an ncu source page (cuda,sass view), SASS rows de-duplicated by address.
usage: python profiles/ncu_phases.py x.csv <kernel substring> name:lo-hi [...]   (line ranges of kernels.cuh)"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
kern = sys.argv[2]
ranges = []
for a in sys.argv[3:]:
    n, r = a.split(":")
    lo, hi = r.split("-")
    ranges.append((n, int(lo), int(hi)))
hdr = next(r for r in rows if r and r[0] == "Line No")
ci = {}
for i, n in enumerate(hdr):
    ci.setdefault(n, i)
seen = {}
fn = sec = None
cur = None
for r in rows:
    if r and r[0] == "File Path":
        sec = r[1].split("/")[-1]
        continue
    if r and r[0] == "Function Name":
        fn = r[1]
        continue
    if len(r) < len(hdr) or r[0] == "Line No":
        continue
    if r[0]:
        cur = (sec, int(r[0]))
        continue
    if fn is None or kern not in fn or r[2] in ("", "..."):
        continue
    try:
        inst = float(r[ci["Instructions Executed"]] or 0)
        smp = float(r[ci["# Samples"]] or 0)
    except ValueError:
        continue
    seen.setdefault(r[2], (cur, inst, smp, r[3].strip()))
agg = {n: [0.0, 0.0, 0] for n, _, _ in ranges}
agg["other"] = [0.0, 0.0, 0]
for addr, ((sec, ln), inst, smp, txt) in seen.items():
    name = "other"
    if sec == "kernels.cuh":
        for n, lo, hi in ranges:
            if lo <= ln <= hi:
                name = n
                break
    agg[name][0] += inst
    agg[name][1] += smp
    agg[name][2] += 1
ti = sum(v[0] for v in agg.values()) or 1
ts = sum(v[1] for v in agg.values()) or 1
print(f"{kern}: {ti:.4e} warp-instructions, {ts:.0f} samples, {len(seen)} SASS instructions")
for n, (i, s, c) in agg.items():
    print(f"{n:12s} {100 * i / ti:5.1f}% inst  {100 * s / ts:5.1f}% samples  ({i:.3e} warp-inst, {c} SASS)")
