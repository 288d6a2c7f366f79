"""Extra pinning run of the CPU oracle against the REFERENCE ICD on scenes that are NOT in the golden set: every scene family
of tests/scenes.py at seeds beyond the committed range (the families draw their geometry from the seed, their state from
seed % k).  Nothing is written; mismatches are listed.  Runs only where oracle/_ref holds the reference build.
usage: python tests/golden/fuzz_pin.py [first_seed] [seeds_per_family]   |   python tests/golden/fuzz_pin.py mixed <first_seed> <count>"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))

import scenes  # noqa: E402
from oracle import swref  # noqa: E402


def mixed(seed: int):
    """One scene with every state drawn independently at random (the families tie state to seed % k): formats, sample count,
    1-3 draws with random compare ops / stencil faces / blend equations / masks / bias / cull / scissor / viewport depth range /
    alphaToCoverage / depth bounds / sampler state, ordinary and special-valued vertices."""
    from scenes import Draw, Scene, StencilFace, Texture  # noqa: F401
    S = scenes
    rng = np.random.default_rng(31000 + seed)
    pick = lambda xs: xs[int(rng.integers(len(xs)))]  # noqa: E731
    samples = pick([1, 1, 4])
    colour = pick([S.FMT_R8G8B8A8_UNORM, S.FMT_B8G8R8A8_UNORM] + ([S.FMT_R8G8B8A8_SRGB, S.FMT_B8G8R8A8_SRGB, S.FMT_R16G16B16A16_SFLOAT, S.FMT_R32G32B32A32_SFLOAT] if samples == 1 else []))
    has_stencil = bool(rng.integers(2))
    depth_fmt = S.FMT_D32_SFLOAT if has_stencil else pick([S.FMT_D32_SFLOAT, S.FMT_D16_UNORM])
    ops = [S.SOP_KEEP, S.SOP_ZERO, S.SOP_REPLACE, S.SOP_INC_CLAMP, S.SOP_DEC_CLAMP, S.SOP_INVERT, S.SOP_INC_WRAP, S.SOP_DEC_WRAP]
    factors = list(range(15))  # VkBlendFactor 0..14
    bops = [S.BOP_ADD, S.BOP_SUBTRACT, S.BOP_REVERSE_SUBTRACT, S.BOP_MIN, S.BOP_MAX]
    draws = []
    for _ in range(int(rng.integers(1, 4))):
        n = int(rng.integers(3, 12))
        tris = []
        for i in range(n):
            v = S._verts(rng, S._tri_kind(rng, int(rng.integers(6))), persp=bool(rng.integers(2)), colour=rng.uniform(-0.2, 1.2, (3, 4)))
            if rng.integers(8) == 0:
                v[int(rng.integers(3)), int(rng.integers(4))] = pick([np.nan, np.inf, -np.inf, 0.0, 1e30, -1e30, 3.4e38, 1e-40])
            tris.append(v)
        face = lambda: StencilFace(failOp=pick(ops), passOp=pick(ops), depthFailOp=pick(ops), compareOp=int(rng.integers(8)),  # noqa: E731
                                   compareMask=pick([0xFF, 0x0F, 0xF3]), writeMask=pick([0xFF, 0x3C, 0x00]), reference=int(rng.integers(256)))
        kw = dict(depthTest=bool(rng.integers(2)), depthWrite=bool(rng.integers(2)), depthCompareOp=int(rng.integers(8)),
                  stencilTest=has_stencil and bool(rng.integers(2)), front=face(), back=face(),
                  blend=bool(rng.integers(2)), srcColor=pick(factors), dstColor=pick(factors), colorOp=pick(bops),
                  srcAlpha=pick(factors), dstAlpha=pick(factors), alphaOp=pick(bops), colorWriteMask=pick([0xF, 0xF, 0x7, 0x5, 0x8]),
                  blendConstants=tuple(float(x) for x in rng.uniform(-0.2, 1.2, 4)), cullMode=pick([S.CULL_NONE, S.CULL_NONE, S.CULL_BACK, S.CULL_FRONT]),
                  frontFace=int(rng.integers(2)), alphaToCoverage=rng.integers(4) == 0, sampleMask=pick([0xF, 0xF, 0x5, 0xA, 0x1]))
        if rng.integers(3) == 0:
            kw["depthBias"] = (float(rng.uniform(-8, 8)), pick([0.0, 0.001, -0.002]), float(rng.uniform(-2, 2)))
        if rng.integers(4) == 0:
            lo = float(np.float32(rng.uniform(0.1, 0.6)))
            kw["depthBounds"] = (lo, float(np.float32(lo + rng.uniform(0.05, 0.4))))
        if rng.integers(3) == 0:
            x, y = int(rng.integers(0, 20)), int(rng.integers(0, 20))
            kw["scissor"] = (x, y, int(rng.integers(8, S.CELL - x + 1)), int(rng.integers(8, S.CELL - y + 1)))
        if rng.integers(4) == 0:
            kw["viewport"] = (float(rng.integers(-8, 8)), float(rng.integers(-8, 8)), float(rng.integers(40, 80)), float(rng.integers(40, 80)),
                              float(rng.uniform(-0.2, 0.4)), float(rng.uniform(0.6, 1.3)))
        fs = "fs_col4"
        if rng.integers(3) == 0:
            w, h = pick([(16, 16), (64, 32), (32, 64)])
            levels = pick([1, 1, int(np.log2(max(w, h))) + 1])
            kw["texture"] = Texture(S._rand_tex(rng, w, h, levels), srgb=bool(rng.integers(2)), magFilter=int(rng.integers(2)), minFilter=int(rng.integers(2)),
                                    mipmapMode=int(rng.integers(2)), addressModeU=pick([S.ADDR_REPEAT, S.ADDR_CLAMP_TO_EDGE, S.ADDR_MIRRORED_REPEAT]),
                                    addressModeV=pick([S.ADDR_REPEAT, S.ADDR_CLAMP_TO_EDGE, S.ADDR_MIRRORED_REPEAT]),
                                    mipLodBias=pick([0.0, 0.75, -0.5]), minLod=pick([0.0, 1.25]), maxLod=float(levels - 1))
            fs = "fs_tex_col4"
            for v in tris:
                v[:, 4:6] = rng.uniform(-4, 5, (3, 2)) * pick([0.1, 0.7, 3.0])
        draws.append(Draw(np.concatenate(tris).astype(np.float32), S.P4C4, "vs_pos4_col4", fs, **kw))
    # the harness' host-side clear does not restate the clear's own sRGB encode / half rounding (scene.py clear_color_bytes)
    if colour in (S.FMT_R8G8B8A8_SRGB, S.FMT_B8G8R8A8_SRGB):
        clear = tuple(float(x) for x in rng.integers(0, 2, 4))
    elif colour == S.FMT_R16G16B16A16_SFLOAT:
        clear = tuple(float(x) / 8.0 for x in rng.integers(0, 9, 4))
    else:
        clear = tuple(float(x) for x in rng.uniform(0, 1, 4))
    return Scene(S.CELL, S.CELL, draws, samples=samples, colorFormat=colour, hasDepth=True, depthFormat=depth_fmt, hasStencil=has_stencil,
                 clearDepth=float(np.float32(rng.uniform(0.3, 1.0))), clearStencil=int(rng.integers(256)), clearColor=clear)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "mixed":
        first, count = int(sys.argv[2]), int(sys.argv[3])
        bad = skipped = 0
        for s in range(first, first + count):
            scene = mixed(s)
            try:
                att = swref.render_oracle(scene)
            except RuntimeError as e:  # a state combination outside the subset (the library rejects it the same way)
                skipped += 1
                continue
            ref = swref.render_reference(scene)
            ref.pop("timing", None)
            res = swref.resolve_oracle(scene, att) if scene.samples > 1 else None
            out = scenes.outputs(scene, att, res)
            for k, v in ref.items():
                nz = int((out[k].view(np.uint8) != v.view(np.uint8)).sum())
                if nz:
                    bad += 1
                    print(f"MISMATCH mixed_{s}/{k}: {nz} bytes", flush=True)
        print(f"mixed scenes={count} outside the subset={skipped} mismatching outputs={bad}")
        return 1 if bad else 0
    # families at seeds outside the golden range

    first = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    bad = total = 0
    for fam, (gen, _) in scenes.FAMILIES.items():
        for s in range(first, first + count):
            try:
                scene = gen(s)
            except Exception as e:  # noqa: BLE001  (a family that indexes a fixed table by seed)
                print(f"{fam}_{s}: generator failed ({e})")
                continue
            ref = swref.render_reference(scene)
            ref.pop("timing", None)
            att = swref.render_oracle(scene)
            res = swref.resolve_oracle(scene, att) if scene.samples > 1 else None
            out = scenes.outputs(scene, att, res)
            total += 1
            for k, v in ref.items():
                nz = int((out[k].view(np.uint8) != v.view(np.uint8)).sum())
                if nz:
                    bad += 1
                    print(f"MISMATCH {fam}_{s}/{k}: {nz} bytes")
        print(fam, "done", flush=True)
    print(f"scenes={total} mismatching outputs={bad}")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
