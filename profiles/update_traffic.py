#!/usr/bin/env python
"""profiles/traffic.json <- dram__bytes_read.sum + dram__bytes_write.sum of the tile kernel in the raw-page exports of one GPU visit
(gpurun_out/<tag>/raw_c4.csv, raw_c5.csv), stamped with the hash of the kernel sources.  Run it BEFORE the kernel sources change again:
bench.py only reports `roofline.traffic` while the stamp matches the sources it runs.   usage: python profiles/update_traffic.py <tag>"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import kernels_hash  # noqa: E402

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
WL = {"c4": ("c4_mesh1m_msaa4_blend_4k", "k_tile<4>"), "c5": ("c5_mesh10m_textured_8k", "k_tile<1>")}


def main():
    tag = sys.argv[1]
    path = os.path.join(ROOT, "profiles", "traffic.json")
    tj = json.load(open(path))
    for short, (name, kernel) in WL.items():
        p = os.path.join(ROOT, "gpurun_out", tag, f"raw_{short}.csv")
        if not os.path.exists(p):
            continue
        rows = list(csv.reader(open(p)))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            if "k_tile" not in r[hdr.index("Kernel Name")]:
                continue
            rd = float(r[hdr.index("dram__bytes_read.sum")]) * UNIT[units[hdr.index("dram__bytes_read.sum")]]
            wr = float(r[hdr.index("dram__bytes_write.sum")]) * UNIT[units[hdr.index("dram__bytes_write.sum")]]
            tj[name] = {"kernel": kernel, "bytes": int(rd + wr), "kernels_sha16": kernels_hash(),
                        "source": f"profiles/{tag}/SUMMARY.md (raw_{short}: {rd / 1e6:.1f} MB read + {wr / 1e6:.1f} MB written)"}
            break
    json.dump(tj, open(path, "w"), indent=2)
    print(json.dumps(tj, indent=2))


if __name__ == "__main__":
    main()
