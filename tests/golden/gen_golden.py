"""Generates tests/golden/* by rendering every case of tests/scenes.py with the REFERENCE ICD
(oracle/_ref/libvk_swiftshader.so, built by oracle/build_ref.sh from /root/reference; runs only in the container
that has the reference).  Run:  python tests/golden/gen_golden.py [--check-oracle]

Outputs (committed):
  golden_hashes.json  — sha256 of every reference output array (colour / depth / stencil) of every case
  golden_full.npz     — the full reference arrays for the first cases of each family (value-level diagnosis)
With --check-oracle the C restatement is rendered beside it and every mismatch is reported (the pinning run).
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))

import scenes  # noqa: E402
from oracle import swref  # noqa: E402

FULL_PER_FAMILY = 2


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    check = "--check-oracle" in sys.argv
    # --only fam1,fam2: render only these families and MERGE them into the committed files (everything else is kept as it is)
    only = None
    if "--only" in sys.argv:
        only = tuple(sys.argv[sys.argv.index("--only") + 1].split(","))
    hashes, full, seen_fam, bad = {}, {}, {}, 0
    if only:
        hashes = json.load(open(os.path.join(HERE, "golden_hashes.json")))
        full = dict(np.load(os.path.join(HERE, "golden_full.npz")))
    for name, scene in scenes.all_cases():
        if only and not name.startswith(only):
            continue
        ref = swref.render_reference(scene)
        ref.pop("timing", None)
        hashes[name] = {k: sha(v) for k, v in ref.items()}
        fam = name.rsplit("_", 1)[0]
        seen_fam[fam] = seen_fam.get(fam, 0) + 1
        if seen_fam[fam] <= FULL_PER_FAMILY and scene.width * scene.height <= 512 * 512:
            for k, v in ref.items():
                full[f"{name}/{k}"] = v
        if check:
            att = swref.render_oracle(scene)
            res = swref.resolve_oracle(scene, att) if scene.samples > 1 else None
            out = scenes.outputs(scene, att, res)
            for k, v in ref.items():
                if not np.array_equal(out[k].view(np.uint8), v.view(np.uint8)):
                    bad += 1
                    if k == "depth":
                        diff = (out[k].view(np.uint32) != v.view(np.uint32)).sum()
                    else:
                        diff = int(np.abs(out[k].astype(int) - v.astype(int)).max())
                    nz = int((out[k].view(np.uint8) != v.view(np.uint8)).sum())
                    print(f"MISMATCH {name}/{k}: bytes differing {nz}, metric {diff}")
        print(name, "ok", flush=True)
    with open(os.path.join(HERE, "golden_hashes.json"), "w") as f:
        json.dump(hashes, f, indent=0, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "golden_full.npz"), **full)
    print(f"cases={len(hashes)} oracle_mismatches={bad}")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
