"""Diagnostic (not a test): renders every golden case with the CUDA path in both modes, compares with the oracle, and
writes gpurun_out/parity_report.json with per-case mismatch counts and the first few mismatching locations."""
import json
import os
import sys
import time
import traceback

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import scenes  # noqa: E402
from oracle import swref  # noqa: E402
from swiftshader_b200.scene import Device  # noqa: E402


def main():
    only = sys.argv[1:] or None
    dev = Device(0)
    report, nbad = {}, 0
    t0 = time.time()
    for name, scene in scenes.all_cases():
        if only and not any(name.startswith(o) for o in only):
            continue
        want = swref.render_oracle(scene)
        for mode in (0, 1):
            key = f"{name}:{'binned' if mode else 'direct'}"
            try:
                dev.set_option("force_binned", mode)
                got = dev.render(scene)
                entry = {}
                for k in want:
                    a, b = got[k], want[k]
                    ne = a.view(np.uint8).reshape(a.shape[0], a.shape[1], a.shape[2], -1) != b.view(np.uint8).reshape(a.shape[0], a.shape[1], a.shape[2], -1)
                    px = np.argwhere(ne.any(axis=-1))
                    if len(px):
                        nbad += 1
                        ex = []
                        for (q, y, x) in px[:6]:
                            ex.append({"q": int(q), "y": int(y), "x": int(x), "got": np.atleast_1d(a[q, y, x]).tolist(), "want": np.atleast_1d(b[q, y, x]).tolist()})
                        entry[k] = {"pixels": int(len(px)), "of": int(a.shape[0] * a.shape[1] * a.shape[2]), "examples": ex}
                report[key] = entry or "ok"
            except Exception as e:  # noqa: BLE001
                nbad += 1
                report[key] = {"error": str(e), "trace": traceback.format_exc()[-800:]}
    dev.set_option("force_binned", 0)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/parity_report.json", "w") as f:
        json.dump(report, f, indent=1)
    ok = sum(1 for v in report.values() if v == "ok")
    print(f"parity: {ok}/{len(report)} ok, {nbad} mismatching outputs, {time.time() - t0:.1f}s")
    fams = {}
    for k, v in report.items():
        fam = k.rsplit("_", 1)[0] + ":" + k.split(":")[1]
        fams.setdefault(fam, [0, 0])
        fams[fam][0 if v == "ok" else 1] += 1
    for fam, (a, b) in sorted(fams.items()):
        print(f"  {fam:32s} ok={a} bad={b}")
    for k, v in list(report.items()):
        if v != "ok":
            print(k, json.dumps(v)[:600])
            break


if __name__ == "__main__":
    main()
