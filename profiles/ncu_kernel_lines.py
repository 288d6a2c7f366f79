#!/usr/bin/env python
"""Per-CUDA-line instruction / stall-sample shares of ONE kernel from an ncu source page (cuda,sass view).
usage: ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > x.csv; python profiles/ncu_kernel_lines.py x.csv <kernel substring> [N]
Only the section of kernels.cuh itself is used (the per-file sections of inlined headers repeat the same SASS rows)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
agg, cur, hdr, ci, active, fname, path = {}, None, None, {}, False, "", ""
seen = set()
for r in rows:
    if r and r[0] == "File Path":
        path = r[1]
        continue
    if r and r[0] == "Function Name":
        fname = r[1]
        continue
    if r and r[0] == "Line No":
        hdr, ci = r, {}
        for i, n in enumerate(hdr):
            ci.setdefault(n, i)
        active = want in fname and path.endswith("kernels.cuh") and (fname, path) not in seen
        seen.add((fname, path))
        cur = None
        continue
    if not active or hdr is None or len(r) < len(hdr):
        continue
    if r[0]:
        cur = (int(r[0]), r[1].strip())
        agg.setdefault(cur, [0.0, 0.0])
        continue
    if cur is None or r[2] in ("", "..."):
        continue
    try:
        agg[cur][0] += float(r[ci["Instructions Executed"]] or 0)
        agg[cur][1] += float(r[ci["# Samples"]] or 0)
    except ValueError:
        pass
ti = sum(a[0] for a in agg.values()) or 1
ts = sum(a[1] for a in agg.values()) or 1
print(f"{want}: {ti:.4e} warp-instructions, {ts:.0f} samples")
for (ln, src), (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{ln:5d} {100 * i / ti:5.1f}%inst {100 * s / ts:5.1f}%smp | {src[:110]}")
