#!/usr/bin/env python
"""Per-phase instruction totals of one kernel from an ncu source page (cuda,sass view), SASS rows de-duplicated by address.
Phases are delimited by marker substrings searched in kernels.cuh, so the tool follows
the code as it moves (the file on disk, or $KERNELS_CUH, must be the profiled version).  usage: python profiles/ncu_phases.py x.csv <kernel substring> [warps]"""
import csv
import sys

MARKERS = [  # (phase, substring of the first line of the phase), in file order
    ("helpers", "small helpers (Reactor semantics"),
    ("vertex+clip", "DEVI uint32_t fetch_index"),
    ("spans", "SetupRoutine::edge (SetupRoutine.cpp:550-621), row-stepping form"),
    ("setup", "DEVI void rot1("),
    ("binning", "Can a fragment of big triangle b land in region"),
    ("sampler", "DEVI uint32_t mulhi16"),
    ("pixhelp", "DEVI bool stencil_compare"),
    ("tile-prolog", "FS (\"fast state\")"),
    ("bin-order", "the bin in triangle order"),
    ("fetch", "32 bin entries at a time, one per lane; the ids"),
    ("cover-big1x", "BIG, 1x: one entry"),
    ("cover-big", "BIG: up to BIG_GROUP"),
    ("cover-small", "SMALL: lanes [p, hi) clip"),
    ("expand", "producer-side expansion: item ="),
    ("queue-wait", "the TMA loads of the region have landed"),
    ("item-fetch", "consume the items 32 at a time"),
    ("item-planes", "plane equations of the item's triangle"),
    ("conflict", "two fragments of this round on one SAMPLE"),
    ("item-shade", "interpolate + routed fragment shader"),
    ("item-tests", "per covered sample: stencil test"),
    ("item-blend", "const uint32_t px = floatTarget"),
    ("item-stencilw", "writeStencil :754-817"),
    ("tile-epilog", "never leave with a bulk copy"),
    ("after", "the steps either side of the draw"),
]
rows = list(csv.reader(open(sys.argv[1])))
kern = sys.argv[2]
warps = float(sys.argv[3]) if len(sys.argv) > 3 else 0
hdr, ci = None, {}
seen = {}
srcline = {}
fn = sec = cur = None
for r in rows:
    if r and r[0] == "File Path":
        sec = r[1].split("/")[-1]
        continue
    if r and r[0] == "Function Name":
        fn = r[1]
        continue
    if r and r[0] == "Line No":  # every section has its own header (the column count varies)
        hdr, ci = r, {}
        for i, n in enumerate(hdr):
            ci.setdefault(n, i)
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[0]:
        cur = (sec, int(r[0]))
        if sec == "kernels.cuh":
            srcline[int(r[0])] = r[1]
        continue
    if fn is None or kern not in fn or r[2] in ("", "..."):
        continue
    try:
        inst = float(r[ci["Instructions Executed"]] or 0)
        smp = float(r[ci["# Samples"]] or 0)
    except ValueError:
        continue
    seen.setdefault(r[2], (cur, inst, smp))
import os
SRC = os.environ.get("KERNELS_CUH", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "swiftshader_b200", "csrc", "kernels.cuh"))
srcline = {i + 1: l for i, l in enumerate(open(SRC).read().splitlines())}  # must be the profiled version of the file
starts = []
for name, sub in MARKERS:
    ln = next((l for l in sorted(srcline) if sub in srcline[l]), None)
    if ln is not None:
        starts.append((ln, name))
starts.sort()
def phase(sec, ln):
    if sec != "kernels.cuh":
        return "intrinsics"
    name = "head"
    for l, n in starts:
        if ln >= l:
            name = n
    return name
agg = {}
for addr, ((sec, ln), inst, smp) in seen.items():
    a = agg.setdefault(phase(sec, ln), [0.0, 0.0, 0])
    a[0] += inst
    a[1] += smp
    a[2] += 1
ti = sum(v[0] for v in agg.values()) or 1
ts = sum(v[1] for v in agg.values()) or 1
print(f"{kern}: {ti:.4e} warp-instructions, {ts:.0f} samples, {len(seen)} SASS instructions")
order = ["head"] + [n for _, n in starts] + ["intrinsics"]
for n in order:
    if n in agg:
        i, s, c = agg[n]
        per = f"  {i / warps:7.0f}/warp" if warps else ""
        print(f"{n:14s} {100 * i / ti:5.1f}% inst  {100 * s / ts:5.1f}% samples  ({i:.3e} warp-inst, {c} SASS){per}")
