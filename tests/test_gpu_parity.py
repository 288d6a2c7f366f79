"""GPU: the CUDA draw path (through the C-ABI, swcu_draw) against the CPU oracle on the same seeded scenes, and against
the committed reference-ICD goldens.  Bars (north_star): coverage / stencil bit-exact, depth <= 1 ULP, UNORM8 colour
<= 1 LSB.  The kernels reproduce the reference's integer and float pipelines operation by operation, so these tests
demand EXACT equality everywhere (0 ULP, 0 LSB) — tighter than the stated tolerance."""
import hashlib
import json
import os

import numpy as np
import pytest

import scenes
from oracle import swref

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
HASHES = json.load(open(os.path.join(HERE, "golden", "golden_hashes.json")))
CASES = dict(scenes.all_cases())


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def cuda_outputs(device, scene):
    """Host-cleared attachments uploaded to the device shadows, every draw through swcu_draw, the end-of-pass resolve on
    the device, everything read back."""
    from swiftshader_b200.scene import Frame
    fr = Frame(device, scene)
    try:
        fr.upload_inputs()
        fr.upload_attachments()
        fr.draw()
        fr.resolve()
        fr.download_all()
        device.sync()
        att = {k: v.copy() for k, v in fr.att.items()}
        res = fr.resolved[0].copy() if fr.resolved is not None else None
    finally:
        fr.close()
    return att, scenes.outputs(scene, att, res)


# direct: no binning; binned: forced through the sort; generic: binned, with the fast-state instantiations of the tile
# kernel switched off, so every scene also runs the instantiation that decodes the fixed-function state at run time
MODES = {"direct": (0, 1), "binned": (1, 1), "generic": (1, 0)}


@pytest.mark.parametrize("mode", sorted(MODES))
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_matches_reference_golden_and_oracle(device, name, mode):
    scene = CASES[name]
    binned, fast = MODES[mode]
    device.set_option("force_binned", binned)
    device.set_option("fast_state", fast)
    try:
        att, out = cuda_outputs(device, scene)
    finally:
        device.set_option("force_binned", 0)
        device.set_option("fast_state", 1)
    for k, h in HASHES[name].items():
        if sha(out[k]) != h:
            # value-level diagnosis against the oracle (which is pinned to the same goldens on the CPU side)
            want = swref.render_oracle(scene)
            wres = swref.resolve_oracle(scene, want) if scene.samples > 1 else None
            wout = scenes.outputs(scene, want, wres)
            bad = np.argwhere(out[k].view(np.uint8) != wout[k].view(np.uint8))
            pytest.fail(f"{name}/{k}: CUDA differs from the reference golden; {len(bad)} bytes differ from the oracle, first at {bad[:5].tolist()}")
    if scene.samples > 1:  # the goldens hold the resolved image; the sample planes are checked against the oracle
        want = swref.render_oracle(scene)
        for k in want:
            assert np.array_equal(att[k].view(np.uint8), want[k].view(np.uint8)), f"{name}/{k}: sample planes differ from the oracle"


MSAA_CASES = sorted(n for n, sc in CASES.items() if sc.samples > 1)


@pytest.mark.parametrize("name", MSAA_CASES)
def test_wide_setup_variant_matches_the_goldens(device, name):
    """k_setup_wide (the 72-register build of the 4x set-up kernel, picked by wave count for a rank's share of a group draw) forced on
    every multisampled golden scene, binned: same bytes."""
    scene = CASES[name]
    device.set_option("force_binned", 1)
    device.set_option("setup_wide", 2)
    try:
        att, out = cuda_outputs(device, scene)
    finally:
        device.set_option("force_binned", 0)
        device.set_option("setup_wide", 1)
    for k, h in HASHES[name].items():
        assert sha(out[k]) == h, f"{name}/{k}: k_setup_wide differs from the reference golden"
