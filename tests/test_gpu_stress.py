"""GPU: overdraw stress — thousands of overlapping triangles of mixed sizes in ONE draw, so that per-pixel order (blending is
not commutative), the conflict handling of the tile kernel (several fragments of one sample in one round), batches that have to
be split because their pairs / items do not fit the shared-memory areas, and the big-triangle / small-triangle paths are all
exercised together.  Compared bit for bit with the CPU oracle (the goldens hold no scene of this density)."""
import numpy as np
import pytest

import scenes
from oracle import swref
from swiftshader_b200.scene import *  # noqa: F401,F403
from swiftshader_b200.scene import Draw, Scene, Texture

pytestmark = pytest.mark.gpu


def _soup(seed: int, n: int, kinds, textured: bool):
    rng = np.random.default_rng(90000 + seed)
    tris = []
    for i in range(n):
        p = scenes._tri_kind(rng, kinds[i % len(kinds)])
        col = rng.uniform(0, 1, (3, 4))
        if textured:
            col[:, :2] = rng.uniform(-2, 3, (3, 2))
        tris.append(scenes._verts(rng, p, persp=(i % 3 == 0), colour=col))
    return np.concatenate(tris, axis=0)


CASES = {
    # name: (samples, n triangles, kinds, textured, depth, blend, W, H)
    "msaa_blend_depth_small": (4, 3000, (1, 4, 1, 5), False, True, True, 160, 96),
    "msaa_blend_mixed": (4, 600, (5, 0, 1, 4), False, False, True, 128, 128),
    "msaa_big_layers": (4, 40, (0, 5), False, True, True, 256, 160),
    "x1_blend_small": (1, 4000, (1, 4, 1), False, False, True, 160, 96),
    "x1_blend_depth_mixed": (1, 800, (5, 0, 1, 4), False, True, True, 192, 128),
    "x1_texture_overdraw": (1, 1500, (1, 5, 4), True, True, False, 128, 96),
    "x1_texture_blend": (1, 300, (0, 5, 1), True, False, True, 96, 96),
}


@pytest.mark.parametrize("fast", [1, 0], ids=["fast", "generic"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_overdraw_soup_matches_oracle(device, name, fast):
    samples, n, kinds, textured, depth, blend, W, H = CASES[name]
    verts = _soup(sorted(CASES).index(name), n, kinds, textured)
    kw = dict(depthTest=depth, depthWrite=depth, blend=blend)
    if textured:
        rng = np.random.default_rng(5)
        kw["texture"] = Texture(scenes._rand_tex(rng, 64, 64, 7), maxLod=6.0)
    d = Draw(verts, scenes.P4C4, "vs_pos4_col4", "fs_tex_col4" if textured else "fs_col4", **kw)
    sc = Scene(W, H, [d], samples=samples, hasDepth=depth, clearDepth=1.0, clearColor=(0.25, 0.5, 0.125, 1.0))
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
        assert len(bad) == 0, f"{name}/{k}: {len(bad)} bytes differ from the oracle, first at {bad[:4].tolist()}"
