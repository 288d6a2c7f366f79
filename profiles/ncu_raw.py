#!/usr/bin/env python
"""Print selected metrics per kernel from an `ncu --page raw --csv` export.  usage: python profiles/ncu_raw.py raw.csv [extra metric substrings...]"""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "launch__grid_size", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_warps", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_config_size", "sm__maximum_warps_per_active_cycle_pct"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
extra = sys.argv[2:]
for r in rows[2:]:
    print("**" + r[hdr.index("Kernel Name")].split("(")[0] + "**", r[hdr.index("Grid Size")], r[hdr.index("Block Size")])
    for i, k in enumerate(hdr):
        if k in KEYS or any(e in k for e in extra):
            print(f"  {k} = {r[i]} {units[i]}")
