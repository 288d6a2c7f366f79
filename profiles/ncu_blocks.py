#!/usr/bin/env python
"""Groups the SASS of ONE kernel (ncu source page, cuda,sass view) into runs of instructions with the same execution count — basic
blocks / loop bodies — and prints the heaviest: n instructions x executions, share of the kernel's warp instructions, stall samples,
and the CUDA lines they belong to.  usage: python profiles/ncu_blocks.py source.csv <kernel substring> [N] [--sass MINEXEC]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 and not sys.argv[3].startswith("--") else 25
path = fname = None
active = seen = False
hdr = None
out = []
for r in rows:
    if r and r[0] == "File Path":
        path = r[1]; continue
    if r and r[0] == "Function Name":
        fname = r[1]; continue
    if r and r[0] == "Line No":
        hdr = r
        ci = {n: i for i, n in reversed(list(enumerate(hdr)))}
        active = (want in fname) and path.endswith("kernels.cuh") and not seen
        seen = seen or active
        cur = None
        continue
    if not active or len(r) < len(hdr):
        continue
    if r[0]:
        cur = int(r[0]); continue
    if r[2] in ("", "..."):
        continue
    try:
        out.append((int(r[2], 16), cur, r[3].strip(), float(r[ci["Instructions Executed"]] or 0), float(r[ci["# Samples"]] or 0)))
    except ValueError:
        pass
out.sort(key=lambda t: t[0])
seenaddr, uniq = set(), []
for t in out:  # a SASS row appears once per CUDA line it is attributed to
    if t[0] not in seenaddr:
        seenaddr.add(t[0]); uniq.append(t)
tot = sum(t[3] for t in uniq) or 1
smp = sum(t[4] for t in uniq) or 1
print(f"{want}: {len(uniq)} SASS instructions, {tot:.4e} warp instructions, {smp:.0f} samples")
runs = []
for a, ln, sass, ex, sm in uniq:
    if runs and abs(runs[-1][2] - ex) <= 1e-9 * max(1.0, ex):
        runs[-1][1] += 1; runs[-1][3] += ex; runs[-1][4] += sm; runs[-1][5].add(ln)
    else:
        runs.append([a, 1, ex, ex, sm, {ln}])
base = uniq[0][0]
for r in sorted(runs, key=lambda r: -r[3])[:top]:
    print(f"+{r[0] - base:#07x} n={r[1]:4d} x {r[2]:.3e} = {100 * r[3] / tot:5.1f}% inst {100 * r[4] / smp:5.1f}% smp  lines {sorted(x for x in r[5] if x)[:9]}")
if "--sass" in sys.argv:
    lim = float(sys.argv[sys.argv.index("--sass") + 1])
    for a, ln, sass, ex, sm in uniq:
        if ex >= lim:
            print(f"+{a - base:#07x} L{ln:<5} {ex:.3e} {sm:6.0f} {sass[:100]}")
