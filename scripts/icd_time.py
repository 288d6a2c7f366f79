#!/usr/bin/env python
"""Times BASELINE workloads THROUGH THE VULKAN API on the patched ICD (oracle/_cuda, the CUDA draw path behind sw::Renderer::draw) and
on the unmodified reference ICD, with the same harness (oracle/refrender --time: vkQueueSubmit -> vkQueueWaitIdle of the draw-only
LOAD pass + resolve).  Not part of bench.py: a measurement of the f2 wiring.   python scripts/icd_time.py [c1 c2 c3 c4 c5]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import swref  # noqa: E402
from swiftshader_b200 import workloads  # noqa: E402

for name in (sys.argv[1:] or ["c1", "c3", "c4"]):
    wl = workloads.WORKLOADS[name]()
    row = {"workload": wl.name}
    frames = 20 if name in ("c1", "c2", "c3") else 8
    t = swref.render_reference(wl.scene, time_frames=frames, warmup=2, icd=swref.CUDA_ICD, env={"SWCU_ICD": "1"})["timing"]
    row["cuda_icd_ms"] = t["median_ms"]
    if swref.reference_available() and name != "c5":
        t = swref.render_reference(wl.scene, time_frames=max(3, frames // 4), warmup=1)["timing"]
        row["reference_icd_ms"] = t["median_ms"]
        row["ratio"] = row["reference_icd_ms"] / row["cuda_icd_ms"]
    print(json.dumps(row), flush=True)
