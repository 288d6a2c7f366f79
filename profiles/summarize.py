#!/usr/bin/env python
"""Turn one GPU-box visit (gpurun_out/<tag>/, written by scripts/gpu_round.sh) into the tracked summary profiles/<tag>/:
bench JSON lines, the ncu launch list (per-kernel totals and shares), the key `ncu --set full` metrics of k_tile / k_setup,
and the hottest CUDA source lines of the tile kernel.   usage: python profiles/summarize.py <tag> [workload ...]"""
import collections
import csv
import glob
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "launch__grid_size", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "smsp__thread_inst_executed_per_inst_executed.ratio"]


def launch_table(path):
    rows = list(csv.reader(open(path)))
    hdr = next(r for r in rows if r and r[0] == "ID")
    agg = collections.OrderedDict()
    for r in rows[rows.index(hdr) + 1:]:
        if len(r) < len(hdr):
            continue
        name = r[hdr.index("Kernel Name")].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[hdr.index("Metric Value")])
    tot = sum(a[1] for a in agg.values()) or 1
    out = ["| kernel | launches | total us | mean us | share |", "|---|---|---|---|---|"]
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{n[:80]}` | {c} | {t / 1e3:.1f} | {t / c / 1e3:.1f} | {100 * t / tot:.1f}% |")
    return "\n".join(out)


def raw_metrics(rawcsv):
    rows = list(csv.reader(open(rawcsv)))
    if len(rows) < 3:
        return "(no rows)"
    hdr, units = rows[0], rows[1]
    out = []
    seen = set()
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0]
        if name in seen:
            continue
        seen.add(name)
        out.append(f"**{name}**  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}")
        out.append("")
        for k in KEYS:
            if k in hdr:
                out.append(f"- `{k}` = {r[hdr.index(k)]} {units[hdr.index(k)]}")
        out.append("")
    return "\n".join(out)


def hot_lines(srccsv, kernel, top=25):
    """the hottest CUDA lines of the first function of the source-page export whose name contains `kernel`"""
    rows = list(csv.reader(open(srccsv)))
    out, keep, fn, taken = [], False, None, False
    pend = []
    for r in rows:
        if r and r[0] == "File Path":
            pend = [r]
            keep = False
            continue
        if r and r[0] == "Function Name":
            fn = r[1]
            keep = (kernel in fn.split("(")[0]) and pend and pend[0][1].endswith("kernels.cuh")
            if keep:
                out += pend
            pend = []
        if keep:
            out.append(r)
    tmp = "/tmp/_src_page.csv"
    with open(tmp, "w", newline="") as f:
        csv.writer(f).writerows(out)
    res = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "ncu_lines.py"), tmp, str(top)], capture_output=True, text=True)
    return res.stdout if res.returncode == 0 else f"(source page unavailable: {res.stderr[-200:]})"


def main():
    tag = sys.argv[1]
    src = os.path.join(ROOT, "gpurun_out", tag)
    dst = os.path.join(ROOT, "profiles", tag)
    os.makedirs(dst, exist_ok=True)
    md = [f"# GPU visit `{tag}` (scripts/gpu_round.sh; one B200)", ""]
    for f in ("smi.txt", "pytest_gpu.log"):
        p = os.path.join(src, f)
        if os.path.exists(p):
            shutil.copy(p, dst)
    p = os.path.join(src, "pytest_gpu.log")
    if os.path.exists(p):
        md += ["## pytest -m gpu", "", "```", "".join(open(p).readlines()[-3:]).strip(), "```", ""]
    md += ["## bench.py (device-timed `value`, end-to-end `e2e`, roofline of the tile kernel, reference ICD on the host cores)", "",
           "| workload | ms/step | Gpix/s | Mtris/s | e2e ms | tile kernel ms | roofline frac (tile kernel) | frame GB/s | reference ICD ms | kernels ms |", "|---|---|---|---|---|---|---|---|---|---|"]
    for p in sorted(glob.glob(os.path.join(src, "bench_c*.json"))):
        shutil.copy(p, dst)
        try:
            j = json.loads(open(p).read().strip().splitlines()[-1])
        except Exception:  # noqa: BLE001
            continue
        r = j["roofline"]
        km = ", ".join(f"{k} {v:.3f}" for k, v in j["kernels_ms"].items())
        cb = j.get("cpu_baseline", {}).get("ms_per_step")
        tile_ms = j["kernels_ms"].get(r["kernel"])
        md.append(f"| {j['config']['workload']} | {j['ms_per_step']:.4f} | {j['value']:.2f} | {j['mtris_per_s']:.1f} | {j['e2e']['ms_per_step']:.3f} | "
                  f"{tile_ms:.4f} | {r['frac']:.4f} ({r['achieved']:.0f} GB/s) | {r['frame_achieved'] or 0:.0f} | {cb if cb is None else round(cb, 2)} | {km} |")
    md.append("")
    for p in sorted(glob.glob(os.path.join(src, "bench_ref_*.json"))):
        shutil.copy(p, dst)
        md += [f"`--impl reference` ({os.path.basename(p)}):", "", "```", open(p).read().strip()[:1200], "```", ""]
    for p in sorted(glob.glob(os.path.join(src, "launches_*.csv"))):
        shutil.copy(p, dst)
        md += [f"## ncu launch list `{os.path.basename(p)}` (cold-cache, serialised: shares only)", "", launch_table(p), ""]
    for p in sorted(glob.glob(os.path.join(src, "raw_c*.csv"))):
        wl = os.path.basename(p)[len("raw_"):-len(".csv")]
        md += [f"## `ncu --set full` of the tile and setup kernels, workload {wl} (raw page)", "", raw_metrics(p), ""]
        srccsv = os.path.join(src, f"source_{wl}.csv")
        if os.path.exists(srccsv):
            warps = {"c4": 64800, "c5": 259200}.get(wl, 0)
            res = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "ncu_phases.py"), srccsv, "k_tile"] + ([str(warps)] if warps else []), capture_output=True, text=True)
            md += ["Tile kernel by phase (warp instructions, stall samples; SASS rows de-duplicated; profiles/ncu_phases.py):", "", "```", res.stdout.strip(), "```", ""]
            md += ["Hottest CUDA lines of the tile kernel (share of warp instructions / of stall samples, top stall reasons):", "", "```",
                   hot_lines(srccsv, "k_tile").strip(), "```", ""]
            md += ["Hottest CUDA lines of k_setup:", "", "```", hot_lines(srccsv, "k_setup", 15).strip(), "```", ""]
    open(os.path.join(dst, "SUMMARY.md"), "w").write("\n".join(md))
    print(f"wrote {dst}/SUMMARY.md")


if __name__ == "__main__":
    main()
