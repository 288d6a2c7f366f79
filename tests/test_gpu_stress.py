"""GPU: overdraw stress (tests/scenes.py: overdraw) — thousands of overlapping triangles of mixed sizes in ONE draw, so that
per-pixel order (blending is not commutative), the conflict handling of the tile kernel (several fragments of one sample in one
round), batches that have to be split because their pairs / items do not fit the shared-memory areas, and the big-triangle /
small-triangle paths are all exercised together.  Every attachment and every sample plane is compared bit for bit with the CPU
oracle, for both instantiations of the tile kernel (the same scenes are also pinned to reference-ICD goldens, test_gpu_parity)."""
import numpy as np
import pytest

import scenes
from oracle import swref

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fast", [1, 0], ids=["fast", "generic"])
@pytest.mark.parametrize("seed", range(7))
def test_overdraw_soup_matches_oracle(device, seed, fast):
    sc = scenes.overdraw(seed)
    device.set_option("force_binned", 1)
    device.set_option("fast_state", fast)
    try:
        got = device.render(sc)
    finally:
        device.set_option("force_binned", 0)
        device.set_option("fast_state", 1)
    want = swref.render_oracle(sc)
    for k in want:
        bad = np.argwhere(got[k].view(np.uint8) != want[k].view(np.uint8))
        assert len(bad) == 0, f"overdraw_{seed}/{k}: {len(bad)} bytes differ from the oracle, first at {bad[:4].tolist()}"
