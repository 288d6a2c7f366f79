"""Generates tests/golden/workload_hashes.json: sha256 of what the UNMODIFIED reference ICD (oracle/_ref) renders for the
five BASELINE.json workloads at FULL size (C1..C5; C4 = the resolved 1x image, C5 = all 4320 rows), plus the reduced-size
instances the multi-GPU check renders.  Runs only where /root/reference has been built into oracle/_ref (this container);
the GPU tests and bench.py compare the CUDA frame with these hashes, so the full-size parity claim rests on the reference
itself and not on the CPU restatement.   python tests/golden/gen_workload_hashes.py [--check-oracle]
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))

from oracle import swref  # noqa: E402
from swiftshader_b200 import workloads  # noqa: E402


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    check = "--check-oracle" in sys.argv
    out, bad = {}, 0
    cases = [(n, workloads.WORKLOADS[n]()) for n in ("c1", "c2", "c3", "c4", "c5")]
    cases += [(f"small_{n}", workloads.small(n)) for n in ("c1", "c2", "c3", "c4", "c5")]
    for key, wl in cases:
        sc = wl.scene
        ref = swref.render_reference(sc)
        ref.pop("timing", None)
        out[key] = {"workload": wl.name, "width": sc.width, "height": sc.height, "samples": sc.samples,
                    "hashes": {k: sha(v) for k, v in ref.items()}}
        if check and not (key == "c5"):  # the oracle needs minutes for 10 M triangles; its row blocks are checked on the GPU side
            att = swref.render_oracle(sc)
            img = swref.resolve_oracle(sc, att)[:sc.height] if sc.samples > 1 else att["color"][0, :sc.height]
            if not np.array_equal(img, ref["color"]):
                bad += 1
                print(f"MISMATCH oracle vs reference: {key}")
        print(key, out[key]["hashes"]["color"][:16], flush=True)
    with open(os.path.join(HERE, "workload_hashes.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(f"cases={len(out)} oracle_mismatches={bad}")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
