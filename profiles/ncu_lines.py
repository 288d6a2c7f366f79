#!/usr/bin/env python
"""Aggregate an ncu source page (cuda,sass view) per CUDA source line.
usage: ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > x.csv; python profiles/ncu_lines.py x.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr, ci, stalls = None, {}, []
cur = None
agg = {}
for r in rows:
    if r and r[0] == "Line No":  # every section of the export has its own header (the column count varies)
        hdr, ci = r, {}
        for i, n in enumerate(hdr):
            ci.setdefault(n, i)
        stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[0]:
        cur = (int(r[0]), r[1].strip())
        agg.setdefault(cur, {"inst": 0.0, "smp": 0.0, "st": {}})
        continue
    if cur is None or r[2] in ("", "..."):
        continue
    a = agg[cur]
    try:
        a["inst"] += float(r[ci["Instructions Executed"]] or 0)
        a["smp"] += float(r[ci["# Samples"]] or 0)
        for s_ in stalls:
            v = float(r[ci[s_]] or 0)
            if v:
                a["st"][s_] = a["st"].get(s_, 0) + v
    except ValueError:
        pass
ti = sum(a["inst"] for a in agg.values()) or 1
ts = sum(a["smp"] for a in agg.values()) or 1
print(f"total warp-instructions {ti:.3e}, samples {ts:.0f}")
tot_st = {}
for a in agg.values():
    for s, v in a["st"].items():
        tot_st[s] = tot_st.get(s, 0) + v
print("stall mix:", ", ".join(f"{s[6:]} {100 * v / ts:.1f}%" for s, v in sorted(tot_st.items(), key=lambda kv: -kv[1])[:8]))
for (ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1]["smp"])[:top]:
    st = ", ".join(f"{s[6:]}:{100 * v / max(a['smp'], 1):.0f}" for s, v in sorted(a["st"].items(), key=lambda kv: -kv[1])[:3])
    print(f"{ln:5d} {100 * a['inst'] / ti:5.1f}%inst {100 * a['smp'] / ts:5.1f}%smp [{st}] | {src[:100]}")
