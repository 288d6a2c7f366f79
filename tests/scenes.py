"""Seeded scene generators shared by the golden generator (tests/golden/gen_golden.py), the CPU tests
(oracle vs reference goldens) and the GPU tests (CUDA vs oracle).  Every family is deterministic in its seed.

The families follow SURVEY.md §8c ("which goldens"): sub-pixel coverage sweeps incl. clipped triangles, cull
modes, the 8 depth compare ops, stencil op matrix, blend-factor matrix, sampler sweeps over filters / address
modes / LODs, 4x MSAA coverage + resolve, index/topology variants, scissor / render-area, and the three
VulkanBenchmarks triangles (tests/VulkanBenchmarks/TriangleBenchmarks.cpp:48-52,100-104,159-163).
"""
from __future__ import annotations

import numpy as np

from swiftshader_b200.scene import *  # noqa: F401,F403
from swiftshader_b200.scene import Draw, Scene, StencilFace, Texture

CELL = 64  # sub-viewport size


def _tri_kind(rng, kind: int):
    if kind == 0:
        p = rng.uniform(-1.6, 1.6, (3, 2))  # big, crosses the frustum sides
    elif kind == 1:
        c = rng.uniform(-0.9, 0.9, 2)
        p = c + rng.uniform(-0.08, 0.08, (3, 2))  # small
    elif kind == 2:
        p = np.round(rng.uniform(-1, 1, (3, 2)) * 32) / 32  # vertices exactly on pixel corners
    elif kind == 3:
        p = (np.round(rng.uniform(-1, 1, (3, 2)) * 32) + 0.5) / 32  # exactly on pixel centres
    elif kind == 4:
        c = rng.uniform(-0.9, 0.9, 2)
        p = c + rng.uniform(-0.02, 0.02, (3, 2))  # tiny / sliver
    else:
        p = rng.uniform(-0.95, 0.95, (3, 2))  # medium, inside
    return p


def _verts(rng, p, persp: bool, colour=None, z=None):
    """(3, 8) float32: clip-space vec4 position + vec4 attribute."""
    w = rng.uniform(0.5, 2.0, 3) if persp else np.ones(3)
    z = rng.uniform(0.1, 0.9, 3) if z is None else z
    v = np.zeros((3, 8), dtype=np.float32)
    v[:, 0] = (p[:, 0] * w).astype(np.float32)
    v[:, 1] = (p[:, 1] * w).astype(np.float32)
    v[:, 2] = (z * w).astype(np.float32)
    v[:, 3] = w.astype(np.float32)
    v[:, 4:8] = 1.0 if colour is None else colour
    return v


P4C4 = [(0, 4, 0), (1, 4, 4)]


def _grid_scene(draw_fn, n=16, **scene_kw) -> Scene:
    """n draws, each confined to its own CELL x CELL viewport+scissor of a square framebuffer."""
    g = int(np.ceil(np.sqrt(n)))
    draws = []
    for i in range(n):
        vx, vy = (i % g) * CELL, (i // g) * CELL
        d = draw_fn(i)
        d.viewport = (float(vx), float(vy), float(CELL), float(CELL), 0.0, 1.0)
        d.scissor = (vx, vy, CELL, CELL)
        draws.append(d)
    return Scene(g * CELL, g * CELL, draws, **scene_kw)


# ------------------------------------------------------------------ families ----
def coverage(seed: int) -> Scene:
    """16 white triangles of mixed kinds; 1/3 with perspective w. Pins R3+R4+R5+R6 coverage incl. the clipper."""
    rng = np.random.default_rng(1000 + seed)

    def mk(i):
        p = _tri_kind(rng, (seed + i) % 5)
        return Draw(_verts(rng, p, persp=(i % 3 == 0)), P4C4, "vs_pos4_col4", "fs_col4")

    return _grid_scene(mk)


def zclip(seed: int) -> Scene:
    """Triangles crossing the near (z<0) and far (z>w) planes, with coloured varyings and depth."""
    rng = np.random.default_rng(2000 + seed)

    def mk(i):
        p = _tri_kind(rng, 5 if i % 2 else 0)
        z = rng.uniform(-0.6, 1.6, 3)
        return Draw(_verts(rng, p, persp=(i % 2 == 0), colour=rng.uniform(0, 1, (3, 4)), z=z), P4C4, "vs_pos4_col4", "fs_col4",
                    depthTest=True, depthWrite=True)

    return _grid_scene(mk, hasDepth=True, clearDepth=1.0)


def zclamp(seed: int) -> Scene:
    """depthClampEnable: no near / far clipping (Context.cpp:647-648), fragment depth clamped to the viewport's depth range instead of
    [0, 1] (PixelProcessor.cpp:121-136) — narrowed and reversed ranges, triangles that cross z < 0 and z > w, two layers per cell."""
    rng = np.random.default_rng(24000 + seed)
    ranges = [(0.2, 0.7), (0.8, 0.3), (0.0, 1.0), (0.45, 0.55)]
    g, draws = 3, []
    for i in range(g * g):
        vx, vy = (i % g) * CELL, (i // g) * CELL
        lo, hi = ranges[(seed + i) % 4]
        for layer in range(2):
            p = _tri_kind(rng, 5 if (i + layer) % 2 else 0)
            z = rng.uniform(-0.6, 1.6, 3)
            d = Draw(_verts(rng, p, persp=(i % 2 == 0), colour=rng.uniform(0, 1, (3, 4)), z=z), P4C4, "vs_pos4_col4", "fs_col4",
                     depthTest=True, depthWrite=True, depthClamp=(layer == 0 or seed % 2 == 0),
                     depthCompareOp=[CMP_LESS_OR_EQUAL, CMP_LESS, CMP_GREATER, CMP_ALWAYS][(seed // 2 + layer) % 4])
            d.viewport = (float(vx), float(vy), float(CELL), float(CELL), lo, hi)
            d.scissor = (vx, vy, CELL, CELL)
            draws.append(d)
    return Scene(g * CELL, g * CELL, draws, samples=4 if seed % 4 == 3 else 1, hasDepth=True, clearDepth=[1.0, 0.5][seed % 2])


def cull(seed: int) -> Scene:
    rng = np.random.default_rng(3000 + seed)
    mode = [CULL_NONE, CULL_FRONT, CULL_BACK, CULL_FRONT | CULL_BACK][seed % 4]
    ff = [FRONT_CCW, FRONT_CW][(seed // 4) % 2]

    def mk(i):
        p = _tri_kind(rng, 5)
        return Draw(_verts(rng, p, persp=(i % 4 == 0)), P4C4, "vs_pos4_col4", "fs_col4", cullMode=mode, frontFace=ff)

    return _grid_scene(mk)


def _layers(rng, n, persp=True, kinds=(5, 0, 1)):
    """n random coloured triangles in ONE draw (tests in-order RMW within a draw)."""
    vs = [_verts(rng, _tri_kind(rng, kinds[i % len(kinds)]), persp and i % 2 == 0, colour=rng.uniform(0, 1, (3, 4))) for i in range(n)]
    return np.concatenate(vs, axis=0)


def depth_ops(seed: int) -> Scene:
    """Two draws of overlapping triangles; second uses compare op seed%8. Pins R7/R8 (+ interpolation R5vi/R6/R11)."""
    rng = np.random.default_rng(4000 + seed)
    op = seed % 8
    d1 = Draw(_layers(rng, 6), P4C4, "vs_pos4_col4", "fs_col4", depthTest=True, depthWrite=True, depthCompareOp=CMP_LESS_OR_EQUAL)
    d2 = Draw(_layers(rng, 6), P4C4, "vs_pos4_col4", "fs_col4", depthTest=True, depthWrite=(seed % 3 != 0), depthCompareOp=op)
    return Scene(CELL, CELL, [d1, d2], hasDepth=True, clearDepth=0.6, clearColor=(0.1, 0.2, 0.3, 1.0))


def depth16(seed: int) -> Scene:
    """D16_UNORM depth buffer (PixelRoutine.cpp:466-482,508-511,687-711): the 8 compare ops on the quantised value, depth
    write with saturating round, the fixed-point constant depth bias (r = 1.01 / 0xFFFF), 4x MSAA, dense overdraw."""
    rng = np.random.default_rng(4500 + seed)
    kw = dict(hasDepth=True, depthFormat=FMT_D16_UNORM, clearColor=(0.1, 0.2, 0.3, 1.0))
    if seed < 8:
        d1 = Draw(_layers(rng, 6), P4C4, "vs_pos4_col4", "fs_col4", depthTest=True, depthWrite=True, depthCompareOp=CMP_LESS_OR_EQUAL)
        d2 = Draw(_layers(rng, 6), P4C4, "vs_pos4_col4", "fs_col4", depthTest=True, depthWrite=(seed % 3 != 0), depthCompareOp=seed % 8)
        return Scene(CELL, CELL, [d1, d2], clearDepth=0.6, **kw)
    if seed == 8:  # constant + slope bias, with and without clamp
        d1 = Draw(_layers(rng, 5), P4C4, "vs_pos4_col4", "fs_col4", depthTest=True, depthWrite=True, depthBias=(3.0, 0.0, 1.5))
        d2 = Draw(_layers(rng, 5), P4C4, "vs_pos4_col4", "fs_col4", depthTest=True, depthWrite=True, depthCompareOp=CMP_LESS, depthBias=(-40.0, -0.0004, 2.0))
        return Scene(CELL, CELL, [d1, d2], clearDepth=1.0, **kw)
    if seed == 9:  # 4x MSAA + blend
        d = Draw(_layers(rng, 10), P4C4, "vs_pos4_col4", "fs_col4", depthTest=True, depthWrite=True, blend=True)
        return Scene(CELL, CELL, [d], samples=4, clearDepth=1.0, **kw)
    if seed == 10:  # many small triangles, binned path
        tris = [_verts(rng, _tri_kind(rng, (1, 4, 5)[i % 3]), persp=(i % 3 == 0), colour=rng.uniform(0, 1, (3, 4))) for i in range(1200)]
        d = Draw(np.concatenate(tris), P4C4, "vs_pos4_col4", "fs_col4", depthTest=True, depthWrite=True, depthCompareOp=CMP_LESS)
        return Scene(128, 96, [d], clearDepth=1.0, **kw)
    # z outside [0,1] through the viewport depth range: clamped before the quantisation
    d = Draw(_layers(rng, 8), P4C4, "vs_pos4_col4", "fs_col4", depthTest=True, depthWrite=True, depthCompareOp=CMP_GREATER,
             viewport=(0.0, 0.0, float(CELL), float(CELL), -0.25, 1.5))
    return Scene(CELL, CELL, [d], clearDepth=0.0, **kw)


def srgb(seed: int) -> Scene:
    """sRGB render targets (R8G8B8A8_SRGB / B8G8R8A8_SRGB): linearToSRGB before the UNORM8 pack (PixelRoutine.cpp:1965-1970),
    sRGBtoLinear of the destination when blending (:1821-1826), both through the relaxed-precision Pow of ShaderCore.cpp.
    Clear colours are 0 / 1 only (their encoding is the identity)."""
    rng = np.random.default_rng(4800 + seed)
    fmt = FMT_B8G8R8A8_SRGB if seed % 3 == 2 else FMT_R8G8B8A8_SRGB
    clear = [(0.0, 0.0, 0.0, 1.0), (1.0, 1.0, 1.0, 0.0), (0.0, 1.0, 0.0, 1.0), (1.0, 0.0, 1.0, 1.0)][seed % 4]
    if seed < 4:  # opaque varying colours: every 8-bit output level of the encode gets hit
        d = Draw(_layers(rng, 8), P4C4, "vs_pos4_col4", "fs_col4", depthTest=(seed % 2 == 1), depthWrite=(seed % 2 == 1))
        return Scene(CELL, CELL, [d], colorFormat=fmt, clearColor=clear, hasDepth=(seed % 2 == 1))
    if seed < 12:  # blending: decode of the destination, blend in linear space, encode
        (sc, dc, co, sa, da, ao) = _BLEND_MATRIX[(seed * 5) % len(_BLEND_MATRIX)]
        d = Draw(_layers(rng, 8), P4C4, "vs_pos4_col4", "fs_col4", blend=True, srcColor=sc, dstColor=dc, colorOp=co,
                 srcAlpha=sa, dstAlpha=da, alphaOp=ao, colorWriteMask=(0xF if seed % 4 else 0xB), blendConstants=(0.25, 0.5, 0.75, 0.4))
        return Scene(CELL, CELL, [d], colorFormat=fmt, clearColor=clear)
    if seed == 12:  # textured
        tex = Texture(_rand_tex(rng, 64, 64, 7), maxLod=6.0)
        tris = []
        for i in range(4):
            col = np.zeros((3, 4))
            col[:, :2] = rng.uniform(-2, 3, (3, 2))
            tris.append(_verts(rng, _tri_kind(rng, [5, 0, 1, 5][i]), persp=(i % 2 == 0), colour=col))
        d = Draw(np.concatenate(tris), P4C4, "vs_pos4_col4", "fs_tex_col4", texture=tex)
        return Scene(CELL, CELL, [d], colorFormat=fmt, clearColor=clear)
    # dense overdraw with SRC_ALPHA blending, binned path
    tris = [_verts(rng, _tri_kind(rng, (1, 4, 5)[i % 3]), persp=(i % 3 == 0), colour=rng.uniform(0, 1, (3, 4))) for i in range(1500)]
    d = Draw(np.concatenate(tris), P4C4, "vs_pos4_col4", "fs_col4", blend=True)
    return Scene(128, 96, [d], colorFormat=fmt, clearColor=clear)


def floatrt(seed: int) -> Scene:
    """Floating-point colour targets, R32G32B32A32_SFLOAT (even seeds) and R16G16B16A16_SFLOAT (odd): no clamping of the shader
    output, the blend factors or the blend constants (PixelProgram.cpp:286-364, PixelRoutine.cpp:1203-1223), the un-folded
    SUBTRACT cases (Context.cpp:1204-1243), destination read / Reactor's Half conversions (Reactor.cpp:3744-3815), masked write."""
    rng = np.random.default_rng(5200 + seed)
    fmt = FMT_R16G16B16A16_SFLOAT if seed % 2 else FMT_R32G32B32A32_SFLOAT
    kw = dict(colorFormat=fmt, clearColor=(0.25, 0.5, 0.125, 1.0))

    def wide(n, kinds=(5, 0, 1)):  # vertex colours outside [0, 1]
        return np.concatenate([_verts(rng, _tri_kind(rng, kinds[i % len(kinds)]), i % 2 == 0, colour=rng.uniform(-0.75, 2.5, (3, 4))) for i in range(n)])
    if seed < 4:
        d = Draw(wide(8), P4C4, "vs_pos4_col4", "fs_col4", depthTest=(seed >= 2), depthWrite=(seed >= 2), colorWriteMask=(0xF if seed < 2 else 0x6))
        return Scene(CELL, CELL, [d], hasDepth=(seed >= 2), **kw)
    if seed < 16:
        (sc, dc, co, sa, da, ao) = _BLEND_MATRIX[(seed * 7 + seed // 2) % len(_BLEND_MATRIX)]
        if seed in (12, 13):  # the two SUBTRACT folds that only apply to UNORM targets
            (sc, dc, co, sa, da, ao) = (BF_ZERO, BF_ONE, BOP_SUBTRACT, BF_SRC_ALPHA, BF_ZERO, BOP_REVERSE_SUBTRACT)
        d = Draw(wide(8), P4C4, "vs_pos4_col4", "fs_col4", blend=True, srcColor=sc, dstColor=dc, colorOp=co, srcAlpha=sa, dstAlpha=da,
                 alphaOp=ao, colorWriteMask=(0xF if seed % 5 else 0xD), blendConstants=(-0.25, 1.5, 0.75, 2.0))
        return Scene(CELL, CELL, [d], **kw)
    if seed < 18:
        tex = Texture(_rand_tex(rng, 64, 64, 7), maxLod=6.0)
        tris = []
        for i in range(4):
            col = np.zeros((3, 4))
            col[:, :2] = rng.uniform(-2, 3, (3, 2))
            tris.append(_verts(rng, _tri_kind(rng, [5, 0, 1, 5][i]), persp=(i % 2 == 0), colour=col))
        d = Draw(np.concatenate(tris), P4C4, "vs_pos4_col4", "fs_tex_col4", texture=tex, blend=(seed == 17))
        return Scene(CELL, CELL, [d], **kw)
    d = Draw(wide(1500, (1, 4, 5)), P4C4, "vs_pos4_col4", "fs_col4", blend=True)  # dense overdraw, binned
    return Scene(128, 96, [d], **kw)


def pathological(seed: int) -> Scene:
    """Vertex data a robust rasteriser must survive exactly like the reference does: NaN / Inf components, w = 0 and w < 0,
    magnitudes of 1e30 and 1e-40, coincident and collinear vertices, zero-area slivers, vertices exactly on the clip planes and
    on the scissor bounds — mixed with ordinary triangles so that order and neighbours matter."""
    rng = np.random.default_rng(6100 + seed)
    # A NaN w is the interesting one: Reactor's CmpNLE is an ordered compare, so the vertex gets no clip flag, projects to the clamp
    # value of RoundIntClamped and the triangle dies in the wrapping row-range arithmetic of the set-up.
    msaa4 = seed % 4 == 3
    specials = [np.nan, np.inf, -np.inf, 0.0, -0.0, 1e30, -1e30, 1e-40, -1e-40, 1.0, -1.0, 1e37, 3.4e38, -3.4e38, -np.nan]
    tris = []
    for i in range(48):
        v = _verts(rng, _tri_kind(rng, (5, 0, 1, 2, 3)[i % 5]), persp=(i % 2 == 0), colour=rng.uniform(0, 1, (3, 4)))
        mode = (i + seed) % 8
        if mode == 0:    # one special value somewhere in a position
            comp, val = rng.integers(4), specials[rng.integers(len(specials))]
            v[rng.integers(3), comp] = val
        elif mode == 1:  # w = 0 / negative w on one or two vertices
            k = rng.integers(3)
            v[k, 3] = [0.0, -0.5, -2.0][rng.integers(3)]
            if rng.integers(2):
                v[(k + 1) % 3, 3] = -1.0
        elif mode == 2:  # coincident / collinear / zero area
            if rng.integers(2):
                v[1, :4] = v[0, :4]
            else:
                v[2, :4] = 0.5 * (v[0, :4] + v[1, :4])
        elif mode == 3:  # exactly on the clip planes: |x| = w, |y| = w, z = 0 / z = w
            k = rng.integers(3)
            v[k, 0] = v[k, 3] * [1.0, -1.0][rng.integers(2)]
            v[(k + 1) % 3, 1] = v[(k + 1) % 3, 3] * [1.0, -1.0][rng.integers(2)]
            v[(k + 2) % 3, 2] = [0.0, v[(k + 2) % 3, 3]][rng.integers(2)]
        elif mode == 4:  # special value in a colour attribute
            v[rng.integers(3), 4 + rng.integers(4)] = specials[rng.integers(len(specials))]
        elif mode == 5:  # far outside the frustum on one side (clipped, huge edge functions)
            v[rng.integers(3), rng.integers(2)] *= 1e6
        elif mode == 6:  # NaN / infinite w on one or two vertices
            k = rng.integers(3)
            v[k, 3] = [np.nan, np.inf, -np.inf, -np.nan][rng.integers(4)]
            if rng.integers(3) == 0:
                v[(k + 1) % 3, 3] = np.nan
        tris.append(v)
    verts = np.concatenate(tris).astype(np.float32)
    d = Draw(verts, P4C4, "vs_pos4_col4", "fs_col4", depthTest=True, depthWrite=True, blend=(seed % 2 == 1),
             scissor=(3, 5, CELL - 7, CELL - 9) if seed % 3 == 0 else None, cullMode=(CULL_NONE, CULL_BACK, CULL_FRONT)[seed % 3])
    return Scene(CELL, CELL, [d], samples=(4 if msaa4 else 1), hasDepth=True, clearDepth=1.0, clearColor=(0.25, 0.5, 0.125, 1.0))


_BLEND_MATRIX = [
    (BF_SRC_ALPHA, BF_ONE_MINUS_SRC_ALPHA, BOP_ADD, BF_ONE, BF_ZERO, BOP_ADD),
    (BF_ONE, BF_ONE, BOP_ADD, BF_ONE, BF_ONE, BOP_ADD),
    (BF_SRC_ALPHA, BF_ONE, BOP_ADD, BF_SRC_ALPHA, BF_ONE_MINUS_SRC_ALPHA, BOP_ADD),
    (BF_DST_COLOR, BF_ZERO, BOP_ADD, BF_DST_ALPHA, BF_ZERO, BOP_ADD),
    (BF_ONE_MINUS_DST_COLOR, BF_SRC_COLOR, BOP_ADD, BF_ONE_MINUS_DST_ALPHA, BF_SRC_ALPHA, BOP_ADD),
    (BF_ONE, BF_ONE_MINUS_SRC_COLOR, BOP_ADD, BF_ONE, BF_ONE_MINUS_SRC_ALPHA, BOP_ADD),
    (BF_SRC_ALPHA_SATURATE, BF_ONE, BOP_ADD, BF_SRC_ALPHA_SATURATE, BF_ONE, BOP_ADD),
    (BF_CONSTANT_COLOR, BF_ONE_MINUS_CONSTANT_COLOR, BOP_ADD, BF_CONSTANT_ALPHA, BF_ONE_MINUS_CONSTANT_ALPHA, BOP_ADD),
    (BF_SRC_ALPHA, BF_ONE_MINUS_SRC_ALPHA, BOP_SUBTRACT, BF_ONE, BF_ONE, BOP_SUBTRACT),
    (BF_SRC_ALPHA, BF_ONE, BOP_REVERSE_SUBTRACT, BF_ONE, BF_ONE, BOP_REVERSE_SUBTRACT),
    (BF_ONE, BF_ONE, BOP_MIN, BF_ONE, BF_ONE, BOP_MIN),
    (BF_ZERO, BF_ZERO, BOP_MAX, BF_SRC_ALPHA, BF_DST_ALPHA, BOP_MAX),
    (BF_ZERO, BF_ONE, BOP_ADD, BF_ONE, BF_ZERO, BOP_ADD),      # colour folds to DST, alpha to SRC
    (BF_ONE, BF_ZERO, BOP_ADD, BF_ZERO, BF_ONE, BOP_ADD),      # colour folds to SRC, alpha to DST
    (BF_ZERO, BF_ZERO, BOP_ADD, BF_ZERO, BF_ZERO, BOP_ADD),    # both fold to ZERO
    (BF_ZERO, BF_ONE, BOP_ADD, BF_ZERO, BF_ONE, BOP_ADD),      # both DST: colour write disabled
    (BF_ZERO, BF_SRC_COLOR, BOP_SUBTRACT, BF_ZERO, BF_ONE, BOP_REVERSE_SUBTRACT),
    (BF_DST_ALPHA, BF_ONE_MINUS_DST_ALPHA, BOP_ADD, BF_ONE_MINUS_SRC_COLOR, BF_DST_COLOR, BOP_ADD),
]


def fragtests(seed: int) -> Scene:
    """alphaToCoverage (PixelRoutine.cpp:643-658 with the thresholds of Renderer.cpp:391-410: the coverage feeds the depth,
    stencil and colour masks, :319-326) and the depth bounds test (:576-641: against the STORED depth, D32F and D16, folding into
    the coverage mask without a depth test and into the depth mask with one, so that the stencil depth-fail op sees it)."""
    rng = np.random.default_rng(4700 + seed)
    col = dict(clearColor=(0.1, 0.2, 0.3, 1.0))

    def layers(n, persp=True):
        v = _layers(rng, n, persp)
        v[:, 7] = rng.uniform(-0.1, 1.1, v.shape[0]).astype(np.float32)  # per-vertex alpha crossing every threshold
        return v

    inc = StencilFace(passOp=SOP_INC_WRAP, failOp=SOP_INVERT, depthFailOp=SOP_DEC_WRAP, compareOp=CMP_ALWAYS)
    if seed < 8:
        ms = 4 if seed % 2 else 1
        kw = {}
        if seed in (2, 3):
            kw = dict(blend=True)
        elif seed in (4, 5):
            kw = dict(depthTest=True, depthWrite=True, depthCompareOp=CMP_LESS_OR_EQUAL, stencilTest=True, front=inc,
                      back=StencilFace(passOp=SOP_REPLACE, depthFailOp=SOP_INC_CLAMP, compareOp=CMP_NOT_EQUAL, reference=3))
        elif seed == 6:
            kw = dict(stencilTest=True, front=inc, back=inc, colorWriteMask=0x7)
        elif seed == 7:
            kw = dict(sampleMask=0b1010, blend=True, depthTest=True, depthWrite=True)
        d = Draw(layers(10), P4C4, "vs_pos4_col4", "fs_col4", alphaToCoverage=True, **kw)
        fmt = dict(colorFormat=FMT_R16G16B16A16_SFLOAT) if seed == 2 else {}
        return Scene(CELL, CELL, [d], samples=ms, hasDepth=True, hasStencil=seed in (4, 5, 6), clearDepth=0.8, **col, **fmt)
    # depth bounds: a first draw lays the stored depth down, the second is bounded by it
    k = seed - 8
    ms = 4 if k % 4 == 3 else 1
    lo = float(np.float32(rng.uniform(0.25, 0.45)))
    hi = float(np.float32(lo + rng.uniform(0.1, 0.3)))
    d1 = Draw(_layers(rng, 8), P4C4, "vs_pos4_col4", "fs_col4", depthTest=True, depthWrite=True, depthCompareOp=CMP_LESS)
    test = k % 2 == 0
    kw = dict(depthTest=test, depthWrite=test and k % 3 != 0, depthCompareOp=(CMP_LESS_OR_EQUAL, CMP_GREATER, CMP_ALWAYS)[k % 3])
    sten = k in (2, 3, 6, 7)
    if sten:
        kw.update(stencilTest=True, front=inc, back=StencilFace(passOp=SOP_REPLACE, depthFailOp=SOP_INVERT, compareOp=CMP_ALWAYS, reference=0x5A))
    d2 = Draw(layers(8), P4C4, "vs_pos4_col4", "fs_col4", depthBounds=(lo, hi), blend=k % 4 == 1, alphaToCoverage=k in (5, 7), **kw)
    fmt = dict(depthFormat=FMT_D16_UNORM) if k in (1, 4, 7) and not sten else {}
    return Scene(CELL, CELL, [d1, d2], samples=ms, hasDepth=True, hasStencil=sten, clearDepth=float(np.float32(rng.uniform(0.3, 0.9))), **col, **fmt)


def blend(seed: int) -> Scene:
    """Blend-factor matrix on RGBA8 with overlapping triangles in one draw (R10/R11 + ordering)."""
    rng = np.random.default_rng(5000 + seed)
    (sc, dc, co, sa, da, ao) = _BLEND_MATRIX[seed % len(_BLEND_MATRIX)]
    wm = 0xF if seed % 5 else 0x5
    d = Draw(_layers(rng, 8), P4C4, "vs_pos4_col4", "fs_col4", blend=True, srcColor=sc, dstColor=dc, colorOp=co,
             srcAlpha=sa, dstAlpha=da, alphaOp=ao, colorWriteMask=wm, blendConstants=(0.25, 0.5, 0.75, 0.4))
    fmt = FMT_B8G8R8A8_UNORM if seed % 7 == 3 else FMT_R8G8B8A8_UNORM
    return Scene(CELL, CELL, [d], clearColor=(0.3, 0.6, 0.2, 0.5), colorFormat=fmt)


def blendoff(seed: int) -> Scene:
    """Blend factors that fold to "keep the destination" while blending is DISABLED: FragmentOutputInterfaceState::colorWriteActive
    (Context.cpp:1304-1308) applies the DST_EXT test to the stored factors whether or not blendEnable is set, so the colour write is
    off (depth still written); plus the neighbouring states that must keep writing."""
    rng = np.random.default_rng(5500 + seed)
    eq = [(BF_ZERO, BF_ONE, BOP_ADD, BF_ZERO, BF_ONE, BOP_ADD),                            # both fold to DST: nothing written
          (BF_ZERO, BF_ONE, BOP_REVERSE_SUBTRACT, BF_ZERO, BF_ONE, BOP_ADD),               # both fold to DST
          (BF_ZERO, BF_ONE, BOP_ADD, BF_ONE, BF_ZERO, BOP_ADD),                            # alpha is SRC: the source is written
          (BF_ZERO, BF_ONE, BOP_ADD, BF_ZERO, BF_ONE, BOP_ADD)][seed % 4]
    (sc, dc, co, sa, da, ao) = eq
    d = Draw(_layers(rng, 6), P4C4, "vs_pos4_col4", "fs_col4", blend=(seed % 4 == 3), srcColor=sc, dstColor=dc, colorOp=co,
             srcAlpha=sa, dstAlpha=da, alphaOp=ao, depthTest=True, depthWrite=True)
    return Scene(CELL, CELL, [d], hasDepth=True, clearColor=(0.3, 0.6, 0.2, 0.5), samples=(4 if seed >= 4 else 1))


def stencil(seed: int) -> Scene:
    """Stencil compare ops x ops, front/back by winding, masks; with and without depth test (R9)."""
    rng = np.random.default_rng(6000 + seed)
    ops = [SOP_KEEP, SOP_ZERO, SOP_REPLACE, SOP_INC_CLAMP, SOP_DEC_CLAMP, SOP_INVERT, SOP_INC_WRAP, SOP_DEC_WRAP]

    def face(k):
        return StencilFace(failOp=ops[(seed + k) % 8], passOp=ops[(seed * 3 + k + 1) % 8], depthFailOp=ops[(seed * 5 + k + 2) % 8],
                           compareOp=(seed + 2 * k) % 8, compareMask=[0xFF, 0x0F, 0xF3][(seed + k) % 3],
                           writeMask=[0xFF, 0xFF, 0x3C][(seed // 2 + k) % 3], reference=(37 * seed + 11 * k + 1) & 0xFF)

    use_depth = seed % 2 == 0
    # first draw lays down a stencil pattern (ALWAYS / INC_WRAP), second exercises the matrix
    d1 = Draw(_layers(rng, 6, persp=False), P4C4, "vs_pos4_col4", "fs_col4", stencilTest=True,
              front=StencilFace(passOp=SOP_INC_WRAP, compareOp=CMP_ALWAYS), back=StencilFace(passOp=SOP_DEC_WRAP, compareOp=CMP_ALWAYS),
              depthTest=use_depth, depthWrite=use_depth)
    d2 = Draw(_layers(rng, 8, persp=False), P4C4, "vs_pos4_col4", "fs_col4", stencilTest=True, front=face(0), back=face(1),
              depthTest=use_depth, depthWrite=use_depth, depthCompareOp=CMP_LESS)
    return Scene(CELL, CELL, [d1, d2], hasDepth=True, hasStencil=True, clearDepth=0.7, clearStencil=(seed * 29) & 0xFF)


def _rand_tex(rng, w, h, levels):
    out = []
    for l in range(levels):
        out.append(rng.integers(0, 256, (max(1, h >> l), max(1, w >> l), 4), dtype=np.uint8))
    return out


def srgbtex(seed: int) -> Scene:
    """The sampler sweep of texture() on an R8G8B8A8_SRGB image: RGB texels go through the reference's sRGBtoLinearFF_FF00 table
    before filtering (SamplerCore.cpp:1966-1977), alpha does not."""
    return texture(seed, srgb=True)


def texture(seed: int, srgb: bool = False) -> Scene:
    """Sampler sweep: uv in [-4,5], filters, mip modes, address modes, LOD bias, independent random mip levels (R14)."""
    rng = np.random.default_rng((7700 if srgb else 7000) + seed)
    variants = [
        dict(w=16, h=16, levels=1, maxLod=0.0),                                   # the benchmark's case
        dict(w=64, h=32, levels=1, maxLod=0.0),
        dict(w=64, h=64, levels=7, maxLod=6.0),                                   # trilinear, full chain
        dict(w=64, h=64, levels=7, maxLod=6.0, mipmapMode=MIPMAP_NEAREST),
        dict(w=32, h=64, levels=4, maxLod=3.0, mipLodBias=0.75),
        dict(w=64, h=64, levels=7, maxLod=6.0, magFilter=FILTER_NEAREST, minFilter=FILTER_NEAREST),
        dict(w=64, h=64, levels=7, maxLod=4.5, minLod=1.25),
        dict(w=32, h=32, levels=6, maxLod=5.0, addressModeU=ADDR_CLAMP_TO_EDGE, addressModeV=ADDR_MIRRORED_REPEAT),
        dict(w=32, h=32, levels=1, maxLod=0.0, addressModeU=ADDR_MIRRORED_REPEAT, addressModeV=ADDR_CLAMP_TO_EDGE),
        dict(w=16, h=16, levels=1, maxLod=0.0, magFilter=FILTER_NEAREST, minFilter=FILTER_NEAREST, mipmapMode=MIPMAP_NEAREST),
    ]
    v = dict(variants[seed % len(variants)])
    tex = Texture(_rand_tex(rng, v.pop("w"), v.pop("h"), v.pop("levels")), srgb=srgb, **v)
    tris = []
    for i in range(3):
        p = _tri_kind(rng, [5, 0, 1][i])
        uvscale = [1.0, 6.0, 30.0][(seed + i) % 3]  # magnified ... strongly minified
        col = np.zeros((3, 4))
        col[:, :2] = rng.uniform(-4, 5, (3, 2)) * (uvscale / 9.0)
        tris.append(_verts(rng, p, persp=(i != 1), colour=col))
    d = Draw(np.concatenate(tris), P4C4, "vs_pos4_col4", "fs_tex_col4", texture=tex)
    return Scene(CELL, CELL, [d])


def texsplit(seed: int) -> Scene:
    """Min and mag filters that differ (FILTER_MIN_POINT_MAG_LINEAR / FILTER_MIN_LINEAR_MAG_POINT, SamplerCore.cpp:278-289): the
    half-texel offset is masked by the sign of the LOD, so the "point" side still runs the 4-tap blend on four coinciding texels -
    on both levels of a trilinear fetch.  Magnified, minified and mixed triangles, every mipmap mode, LOD bias and clamps."""
    rng = np.random.default_rng(7300 + seed)
    mag, mn = ((FILTER_LINEAR, FILTER_NEAREST), (FILTER_NEAREST, FILTER_LINEAR))[seed % 2]
    variants = [
        dict(w=64, h=64, levels=7, maxLod=6.0),
        dict(w=64, h=64, levels=7, maxLod=6.0, mipmapMode=MIPMAP_NEAREST),
        dict(w=32, h=64, levels=4, maxLod=3.0, mipLodBias=0.75),
        dict(w=64, h=32, levels=1, maxLod=0.0),
        dict(w=64, h=64, levels=7, maxLod=4.5, minLod=1.25, addressModeU=ADDR_CLAMP_TO_EDGE, addressModeV=ADDR_MIRRORED_REPEAT),
        dict(w=32, h=32, levels=6, maxLod=5.0, mipLodBias=-0.5),
    ]
    v = dict(variants[(seed // 2) % len(variants)])
    tex = Texture(_rand_tex(rng, v.pop("w"), v.pop("h"), v.pop("levels")), srgb=(seed % 3 == 2), magFilter=mag, minFilter=mn, **v)
    tris = []
    for i in range(4):
        p = _tri_kind(rng, [5, 0, 1, 5][i])
        uvscale = [1.0, 6.0, 30.0, 0.3][(seed + i) % 4]
        col = np.zeros((3, 4))
        col[:, :2] = rng.uniform(-4, 5, (3, 2)) * (uvscale / 9.0)
        tris.append(_verts(rng, p, persp=(i % 2 == 0), colour=col))
    d = Draw(np.concatenate(tris), P4C4, "vs_pos4_col4", "fs_tex_col4", texture=tex)
    return Scene(CELL, CELL, [d], samples=(4 if seed % 5 == 4 else 1))


def msaa(seed: int) -> Scene:
    """4x MSAA: per-sample coverage (R5 iv), resolve; odd seeds add depth test + blending."""
    rng = np.random.default_rng(8000 + seed)
    if seed % 2 == 0:
        def mk(i):
            kind = (seed // 2 + i) % 5
            p = _tri_kind(rng, kind)
            if i % 4 == 3:
                p = np.round(p * 32 * 8) / (32 * 8)  # vertices on the 1/8-pixel sample grid
            return Draw(_verts(rng, p, persp=(i % 3 == 0)), P4C4, "vs_pos4_col4", "fs_col4")
        return _grid_scene(mk, samples=4)
    d = Draw(_layers(rng, 10), P4C4, "vs_pos4_col4", "fs_col4", depthTest=True, depthWrite=True, blend=True)
    return Scene(CELL, CELL, [d], samples=4, hasDepth=True, clearDepth=1.0, clearColor=(0.2, 0.2, 0.2, 1.0))


def topology(seed: int) -> Scene:
    """Indexed (u16/u32) lists, strips and fans, firstIndex / vertexOffset (R2 setBatchIndices)."""
    rng = np.random.default_rng(9000 + seed)
    n = 24
    pts = rng.uniform(-0.95, 0.95, (n, 2))
    v = np.zeros((n, 8), dtype=np.float32)
    v[:, 0:2] = pts
    v[:, 2] = rng.uniform(0.1, 0.9, n)
    v[:, 3] = 1.0
    v[:, 4:8] = rng.uniform(0, 1, (n, 4))
    kind = seed % 6
    kw = {}
    if kind == 0:
        idx, topo = rng.integers(0, n, 30).astype(np.uint16), TOPO_TRIANGLE_LIST
    elif kind == 1:
        idx, topo = rng.integers(0, n - 4, 33).astype(np.uint32), TOPO_TRIANGLE_LIST
        kw = dict(first=3, count=27, vertexOffset=4)
    elif kind == 2:
        idx, topo = None, TOPO_TRIANGLE_STRIP
    elif kind == 3:
        idx, topo = None, TOPO_TRIANGLE_FAN
    elif kind == 4:
        idx, topo = rng.integers(0, n, 12).astype(np.uint16), TOPO_TRIANGLE_STRIP
    else:
        idx, topo = None, TOPO_TRIANGLE_LIST
        kw = dict(first=3, count=18)
    # flat-ish fans/strips get culled differently: exercise cull with strips' alternating winding
    d = Draw(v, P4C4, "vs_pos4_col4", "fs_col4", indices=idx, topology=topo, cullMode=[CULL_NONE, CULL_BACK][seed % 2],
             depthTest=True, depthWrite=True, **kw)
    return Scene(CELL, CELL, [d], hasDepth=True)


def scissor(seed: int) -> Scene:
    """Scissor smaller than the viewport, viewport offset/non-square, depth range != [0,1], depth bias."""
    rng = np.random.default_rng(10000 + seed)
    W, H = 96, 80
    vp = [(0.0, 0.0, 96.0, 80.0, 0.0, 1.0), (8.0, 4.0, 70.0, 60.0, 0.2, 0.9), (-10.0, -6.0, 120.0, 100.0, 0.0, 1.0), (5.5, 3.25, 64.5, 48.75, 1.0, 0.0)][seed % 4]
    sc = [(10, 7, 50, 41), (0, 0, 96, 80), (33, 20, 17, 55), (1, 1, 94, 78)][(seed // 2) % 4]
    bias = [(0.0, 0.0, 0.0), (2.0, 0.0, 1.5), (-3.0, -1e-6, 0.0), (4.0, 1e-7, 2.0)][seed % 4]
    d1 = Draw(_layers(rng, 8), P4C4, "vs_pos4_col4", "fs_col4", viewport=vp, scissor=sc, depthTest=True, depthWrite=True, depthBias=bias,
              depthCompareOp=CMP_LESS if vp[4] <= vp[5] else CMP_GREATER)
    return Scene(W, H, [d1], hasDepth=True, clearDepth=1.0 if vp[4] <= vp[5] else 0.0)


_OVERDRAW = [
    # samples, n triangles, kinds, textured, depth, blend, W, H
    (4, 3000, (1, 4, 1, 5), False, True, True, 160, 96),
    (4, 600, (5, 0, 1, 4), False, False, True, 128, 128),
    (4, 40, (0, 5), False, True, True, 256, 160),
    (1, 4000, (1, 4, 1), False, False, True, 160, 96),
    (1, 800, (5, 0, 1, 4), False, True, True, 192, 128),
    (1, 1500, (1, 5, 4), True, True, False, 128, 96),
    (1, 300, (0, 5, 1), True, False, True, 96, 96),
]


def overdraw(seed: int) -> Scene:
    """Thousands of overlapping triangles of mixed sizes in ONE draw: per-pixel order under non-commutative blending, many
    fragments per sample, big and pixel-sized triangles together (the density the other families do not reach)."""
    samples, n, kinds, textured, depth, blend_on, W, H = _OVERDRAW[seed % len(_OVERDRAW)]
    rng = np.random.default_rng(90000 + seed)
    tris = []
    for i in range(n):
        p = _tri_kind(rng, kinds[i % len(kinds)])
        col = rng.uniform(0, 1, (3, 4))
        if textured:
            col[:, :2] = rng.uniform(-2, 3, (3, 2))
        tris.append(_verts(rng, p, persp=(i % 3 == 0), colour=col))
    kw = dict(depthTest=depth, depthWrite=depth, blend=blend_on)
    if textured:
        kw["texture"] = Texture(_rand_tex(np.random.default_rng(5), 64, 64, 7), maxLod=6.0)
    d = Draw(np.concatenate(tris, axis=0), P4C4, "vs_pos4_col4", "fs_tex_col4" if textured else "fs_col4", **kw)
    return Scene(W, H, [d], samples=samples, hasDepth=depth, clearDepth=1.0, clearColor=(0.25, 0.5, 0.125, 1.0))


def _checker16():
    """The benchmark's 16x16 checkerboard (TriangleBenchmarks.cpp:219-241)."""
    rgb = [0xFFFF0000, 0xFF00FF00, 0xFF0000FF]
    data = np.zeros(256, dtype=np.uint32)
    k = 0
    for i in range(16):
        for j in range(16):
            if ((i ^ j) & 1) == 0:
                data[i + 16 * j] = rgb[k % 3]
                k += 1
    return data.view(np.uint8).reshape(16, 16, 4)


def benchmark(which: int, width=1280, height=720, samples=1) -> Scene:
    """The three VulkanBenchmarks triangles. Defaults mirror the reference harness (1280x720 swapchain,
    tests/VulkanWrapper/DrawTester.hpp:135; clear (0.5,0.5,0.5,1), DrawTester.cpp:325-406; no depth attachment)."""
    clear = (0.5, 0.5, 0.5, 1.0)
    if which == 0:  # TriangleSolidColor
        v = np.array([[1, 1, .5], [-1, 1, .5], [0, -1, .5]], dtype=np.float32)
        d = Draw(v, [(0, 3, 0)], "vs_pos3", "fs_white")
    elif which == 1:  # TriangleInterpolateColor
        v = np.array([[1, 1, .05, 1, 0, 0], [-1, 1, .5, 0, 1, 0], [0, -1, .5, 0, 0, 1]], dtype=np.float32)
        d = Draw(v, [(0, 3, 0), (1, 3, 3)], "vs_pos3_col3", "fs_col3")
    else:  # TriangleSampleTexture
        v = np.array([[1, 1, .5, 1, 0], [-1, 1, .5, 0, 1], [0, -1, .5, 0, 0]], dtype=np.float32)
        d = Draw(v, [(0, 3, 0), (1, 2, 3)], "vs_pos3_uv2", "fs_tex_uv2", texture=Texture([_checker16()], maxLod=0.0))
    return Scene(width, height, [d], samples=samples, clearColor=clear)


def mixed(seed: int):
    """One scene with every state drawn independently at random (the families tie state to seed % k): formats, sample count,
    1-3 draws with random compare ops / stencil faces / blend equations / masks / bias / cull / scissor / viewport depth range /
    alphaToCoverage / depth bounds / sampler state, ordinary and special-valued vertices."""
    rng = np.random.default_rng(31000 + seed)
    pick = lambda xs: xs[int(rng.integers(len(xs)))]  # noqa: E731
    samples = pick([1, 1, 4])
    colour = pick([FMT_R8G8B8A8_UNORM, FMT_B8G8R8A8_UNORM] + ([FMT_R8G8B8A8_SRGB, FMT_B8G8R8A8_SRGB, FMT_R16G16B16A16_SFLOAT, FMT_R32G32B32A32_SFLOAT] if samples == 1 else []))
    has_stencil = bool(rng.integers(2))
    depth_fmt = FMT_D32_SFLOAT if has_stencil else pick([FMT_D32_SFLOAT, FMT_D16_UNORM])
    ops = [SOP_KEEP, SOP_ZERO, SOP_REPLACE, SOP_INC_CLAMP, SOP_DEC_CLAMP, SOP_INVERT, SOP_INC_WRAP, SOP_DEC_WRAP]
    factors = list(range(15))  # VkBlendFactor 0..14
    bops = [BOP_ADD, BOP_SUBTRACT, BOP_REVERSE_SUBTRACT, BOP_MIN, BOP_MAX]
    draws = []
    for _ in range(int(rng.integers(1, 4))):
        n = int(rng.integers(3, 12))
        tris = []
        for i in range(n):
            v = _verts(rng, _tri_kind(rng, int(rng.integers(6))), persp=bool(rng.integers(2)), colour=rng.uniform(-0.2, 1.2, (3, 4)))
            if rng.integers(8) == 0:
                v[int(rng.integers(3)), int(rng.integers(4))] = pick([np.nan, np.inf, -np.inf, 0.0, 1e30, -1e30, 3.4e38, 1e-40])
            tris.append(v)
        face = lambda: StencilFace(failOp=pick(ops), passOp=pick(ops), depthFailOp=pick(ops), compareOp=int(rng.integers(8)),  # noqa: E731
                                   compareMask=pick([0xFF, 0x0F, 0xF3]), writeMask=pick([0xFF, 0x3C, 0x00]), reference=int(rng.integers(256)))
        kw = dict(depthTest=bool(rng.integers(2)), depthWrite=bool(rng.integers(2)), depthCompareOp=int(rng.integers(8)),
                  stencilTest=has_stencil and bool(rng.integers(2)), front=face(), back=face(),
                  blend=bool(rng.integers(2)), srcColor=pick(factors), dstColor=pick(factors), colorOp=pick(bops),
                  srcAlpha=pick(factors), dstAlpha=pick(factors), alphaOp=pick(bops), colorWriteMask=pick([0xF, 0xF, 0x7, 0x5, 0x8]),
                  blendConstants=tuple(float(x) for x in rng.uniform(-0.2, 1.2, 4)), cullMode=pick([CULL_NONE, CULL_NONE, CULL_BACK, CULL_FRONT]),
                  frontFace=int(rng.integers(2)), alphaToCoverage=rng.integers(4) == 0, sampleMask=pick([0xF, 0xF, 0x5, 0xA, 0x1]))
        if rng.integers(3) == 0:
            kw["depthBias"] = (float(rng.uniform(-8, 8)), pick([0.0, 0.001, -0.002]), float(rng.uniform(-2, 2)))
        if rng.integers(4) == 0:
            lo = float(np.float32(rng.uniform(0.1, 0.6)))
            kw["depthBounds"] = (lo, float(np.float32(lo + rng.uniform(0.05, 0.4))))
        if rng.integers(3) == 0:
            x, y = int(rng.integers(0, 20)), int(rng.integers(0, 20))
            kw["scissor"] = (x, y, int(rng.integers(8, CELL - x + 1)), int(rng.integers(8, CELL - y + 1)))
        if rng.integers(4) == 0:
            kw["viewport"] = (float(rng.integers(-8, 8)), float(rng.integers(-8, 8)), float(rng.integers(40, 80)), float(rng.integers(40, 80)),
                              float(rng.uniform(-0.2, 0.4)), float(rng.uniform(0.6, 1.3)))
        fs = "fs_col4"
        if rng.integers(3) == 0:
            w, h = pick([(16, 16), (64, 32), (32, 64)])
            levels = pick([1, 1, int(np.log2(max(w, h))) + 1])
            kw["texture"] = Texture(_rand_tex(rng, w, h, levels), srgb=bool(rng.integers(2)), magFilter=int(rng.integers(2)), minFilter=int(rng.integers(2)),
                                    mipmapMode=int(rng.integers(2)), addressModeU=pick([ADDR_REPEAT, ADDR_CLAMP_TO_EDGE, ADDR_MIRRORED_REPEAT]),
                                    addressModeV=pick([ADDR_REPEAT, ADDR_CLAMP_TO_EDGE, ADDR_MIRRORED_REPEAT]),
                                    mipLodBias=pick([0.0, 0.75, -0.5]), minLod=pick([0.0, 1.25]), maxLod=float(levels - 1))
            fs = "fs_tex_col4"
            for v in tris:
                v[:, 4:6] = rng.uniform(-4, 5, (3, 2)) * pick([0.1, 0.7, 3.0])
        verts, attribs, vs = np.concatenate(tris).astype(np.float32), P4C4, "vs_pos4_col4"
        layout = pick(["p4c4", "p4c4", "p3c3", "p3uv2", "p3"])
        if layout != "p4c4":  # the reference benchmarks' vertex layouts: vec3 position (w = 1), vec3 colour / vec2 uv / nothing
            with np.errstate(all="ignore"):
                pos = verts[:, :3] / np.where(np.isfinite(verts[:, 3:4]) & (verts[:, 3:4] != 0), verts[:, 3:4], 1.0)
            if layout == "p3c3":
                verts, attribs, vs, fs = np.concatenate([pos, verts[:, 4:7]], axis=1), [(0, 3, 0), (1, 3, 3)], "vs_pos3_col3", "fs_col3"
                kw.pop("texture", None)
            elif layout == "p3uv2" and "texture" in kw:
                verts, attribs, vs, fs = np.concatenate([pos, verts[:, 4:6]], axis=1), [(0, 3, 0), (1, 2, 3)], "vs_pos3_uv2", "fs_tex_uv2"
            elif layout == "p3":
                verts, attribs, vs, fs = pos, [(0, 3, 0)], "vs_pos3", "fs_white"
                kw.pop("texture", None)
            verts = np.ascontiguousarray(verts, dtype=np.float32)
        nv = verts.shape[0]
        topo = pick([TOPO_TRIANGLE_LIST, TOPO_TRIANGLE_LIST, TOPO_TRIANGLE_STRIP, TOPO_TRIANGLE_FAN])
        if rng.integers(3) == 0:
            kw["indices"] = rng.integers(0, nv, int(rng.integers(3, 40))).astype(pick([np.uint16, np.uint32]))
            if rng.integers(2) and len(kw["indices"]) >= 7:  # a sub-range that stays inside the index buffer
                kw["first"] = int(rng.integers(0, 4))
                kw["count"] = len(kw["indices"]) - kw["first"] - int(rng.integers(0, 2))
        elif rng.integers(3) == 0 and nv >= 7:
            kw["first"] = int(rng.integers(0, 4))
            kw["count"] = nv - kw["first"] - int(rng.integers(0, 2))
        draws.append(Draw(verts, attribs, vs, fs, topology=topo, **kw))
    # the harness' host-side clear does not restate the clear's own sRGB encode / half rounding (scene.py clear_color_bytes)
    if colour in (FMT_R8G8B8A8_SRGB, FMT_B8G8R8A8_SRGB):
        clear = tuple(float(x) for x in rng.integers(0, 2, 4))
    elif colour == FMT_R16G16B16A16_SFLOAT:
        clear = tuple(float(x) / 8.0 for x in rng.integers(0, 9, 4))
    else:
        clear = tuple(float(x) for x in rng.uniform(0, 1, 4))
    return Scene(CELL, CELL, draws, samples=samples, colorFormat=colour, hasDepth=True, depthFormat=depth_fmt, hasStencil=has_stencil,
                 clearDepth=float(np.float32(rng.uniform(0.3, 1.0))), clearStencil=int(rng.integers(256)), clearColor=clear)



# ------------------------------------------------------------------ f4: a transform in the vertex stage, lines, points ----
P3C4 = [(0, 3, 0), (1, 4, 3)]


def _perspective(rng):
    """A model-view-projection matrix (row-major numpy; the block stores it column-major like a GLSL mat4)."""
    a, b, c = rng.uniform(-0.6, 0.6, 3)
    ca, sa, cb, sb, cc, sc = np.cos(a), np.sin(a), np.cos(b), np.sin(b), np.cos(c), np.sin(c)
    rx = np.array([[1, 0, 0, 0], [0, ca, -sa, 0], [0, sa, ca, 0], [0, 0, 0, 1]])
    ry = np.array([[cb, 0, sb, 0], [0, 1, 0, 0], [-sb, 0, cb, 0], [0, 0, 0, 1]])
    rz = np.array([[cc, -sc, 0, 0], [sc, cc, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]])
    t = np.eye(4)
    t[:3, 3] = (rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3), rng.uniform(2.0, 3.5))
    f, n_, fa = 1.0 / np.tan(rng.uniform(0.5, 0.9) / 2), 0.5, 8.0
    proj = np.array([[f, 0, 0, 0], [0, f, 0, 0], [0, 0, fa / (fa - n_), -fa * n_ / (fa - n_)], [0, 0, 1, 0]])
    return (proj @ t @ rz @ ry @ rx).astype(np.float32)


def mvp(seed: int) -> Scene:
    """A transform in the vertex stage from the push-constant block: OpMatrixTimesVector (MulAdd accumulation) on a tessellated
    object-space patch, or component-wise scale / offset / tint; depth-tested, some of it clipped by the frustum."""
    rng = np.random.default_rng(21000 + seed)
    g = 6 + seed % 5
    u, v = np.meshgrid(np.linspace(-1, 1, g + 1), np.linspace(-1, 1, g + 1))
    z = 0.35 * np.sin(2.5 * u + seed) * np.cos(2.0 * v)
    verts = np.zeros(((g + 1) * (g + 1), 7), dtype=np.float32)
    verts[:, 0], verts[:, 1], verts[:, 2] = u.ravel(), v.ravel(), z.ravel()
    verts[:, 3:7] = rng.uniform(0, 1, (len(verts), 4))
    idx = []
    for j in range(g):
        for i in range(g):
            a = j * (g + 1) + i
            idx += [a, a + 1, a + g + 1, a + 1, a + g + 2, a + g + 1]
    idx = np.array(idx, dtype=np.uint16 if seed % 2 else np.uint32)
    if seed % 3 == 2:
        pc = np.concatenate([rng.uniform(0.4, 1.3, 3), [0.0], rng.uniform(-0.4, 0.4, 2), [rng.uniform(0.2, 0.8)], [rng.uniform(0.8, 1.6)],
                             rng.uniform(0.3, 1.0, 4)]).astype(np.float32)
        vs = "vs_xform_pos3_col4"
    else:
        m = _perspective(rng)
        if seed % 4 == 1:
            m = m * np.float32(rng.uniform(0.8, 1.9))  # pushes part of the patch across the frustum planes
        pc = np.ascontiguousarray(m.T).ravel()
        vs = "vs_mvp_pos3_col4"
    d = Draw(verts, P3C4, vs, "fs_col4", indices=idx, depthTest=True, depthWrite=True, pushConstants=pc,
             cullMode=[CULL_NONE, CULL_FRONT, CULL_NONE][seed % 3], blend=(seed % 4 == 3))
    return Scene(2 * CELL, 2 * CELL, [d], samples=4 if seed % 5 == 4 else 1, hasDepth=True, clearColor=(0.1, 0.1, 0.1, 1.0))


def ubo(seed: int) -> Scene:
    """A transform in the vertex stage from a UNIFORM BUFFER (set 0, binding 1): a column-major model matrix, a row-major
    view-projection matrix (two OpMatrixTimesVector, 32 MulAdd steps) and a tint; 1x and 4x, depth-tested, culled, blended, part of
    the patch across the frustum planes; with a texture in the fragment stage the set holds both descriptors."""
    rng = np.random.default_rng(27000 + seed)
    g = 5 + seed % 4
    u, v = np.meshgrid(np.linspace(-1, 1, g + 1), np.linspace(-1, 1, g + 1))
    z = 0.3 * np.cos(2.0 * u - seed) * np.sin(2.5 * v + 0.5 * seed)
    textured = seed % 4 == 2
    verts = np.zeros(((g + 1) * (g + 1), 7), dtype=np.float32)
    verts[:, 0], verts[:, 1], verts[:, 2] = u.ravel(), v.ravel(), z.ravel()
    verts[:, 3:7] = rng.uniform(0, 1, (len(verts), 4))
    idx = []
    for j in range(g):
        for i in range(g):
            a = j * (g + 1) + i
            idx += [a, a + 1, a + g + 1, a + 1, a + g + 2, a + g + 1]
    idx = np.array(idx, dtype=np.uint32 if seed % 2 else np.uint16)
    model = np.eye(4)
    model[:3, :3] *= rng.uniform(0.45, 0.95, 3)
    model[:3, 3] = rng.uniform(-0.2, 0.2, 3)
    vp = _perspective(rng)
    if seed % 3 == 1:
        vp = vp * np.float32(rng.uniform(0.9, 1.8))  # part of the patch leaves the frustum
    words = np.concatenate([np.ascontiguousarray(model.astype(np.float32).T).ravel(),  # column-major
                            np.ascontiguousarray(vp.astype(np.float32)).ravel(),           # row-major
                            rng.uniform(0.3, 1.0, 4).astype(np.float32),
                            np.zeros(seed % 3 * 4, dtype=np.float32)])                      # (a buffer longer than the block)
    kw = {}
    fs = "fs_col4"
    if textured:
        fs = "fs_tex_col4"
        kw["texture"] = Texture(_rand_tex(rng, 32, 32, 4), maxLod=3.0)
    d = Draw(verts, P3C4, "vs_ubo_pos3_col4", fs, indices=idx, depthTest=True, depthWrite=True, uniformBuffer=(0, 1, words),
             cullMode=[CULL_NONE, CULL_FRONT, CULL_NONE][seed % 3], blend=(seed % 4 == 3), **kw)
    return Scene(2 * CELL, 2 * CELL, [d], samples=4 if seed % 5 == 3 else 1, hasDepth=True, clearColor=(0.05, 0.1, 0.15, 1.0))


def lines(seed: int) -> Scene:
    """Line lists and strips (DrawCall::setupLine, Renderer.cpp:920-1000: the rectangle of the default rasterization mode), with
    perspective, end points outside the frustum, u16 indices, 1x and 4x, with and without a depth test."""
    rng = np.random.default_rng(22000 + seed)
    n = 24
    w = rng.uniform(0.5, 2.0, n) if seed % 2 else np.ones(n)
    span = 1.5 if seed % 3 == 0 else 0.95
    v = np.zeros((n, 8), dtype=np.float32)
    v[:, 0] = rng.uniform(-span, span, n) * w
    v[:, 1] = rng.uniform(-span, span, n) * w
    v[:, 2] = rng.uniform(0.1, 0.9, n) * w
    v[:, 3] = w
    v[:, 4:8] = rng.uniform(0, 1, (n, 4))
    if seed % 8 == 5:
        v[3, 3] = -0.4  # an end point behind the eye
        v[7, 0:2] = v[6, 0:2] / v[6, 3] * v[7, 3]  # a segment of zero length on screen
    kind = seed % 4
    kw = {}
    if kind == 0:
        idx, topo = None, TOPO_LINE_LIST
    elif kind == 1:
        idx, topo = None, TOPO_LINE_STRIP
    elif kind == 2:
        idx, topo = rng.integers(0, n, 20).astype(np.uint16), TOPO_LINE_LIST
    else:
        idx, topo = rng.integers(0, n, 14).astype(np.uint32), TOPO_LINE_STRIP
        kw = dict(first=2, count=11)
    d = Draw(v, P4C4, "vs_pos4_col4", "fs_col4", indices=idx, topology=topo, depthTest=(seed % 3 == 1), depthWrite=True,
             lineWidth=[1.0, 1.0, 3.0, 1.0, 2.5][seed % 5], depthBias=(2.0, 0.0, 1.5) if seed % 6 == 4 else (0.0, 0.0, 0.0), **kw)
    return Scene(2 * CELL, 2 * CELL, [d], samples=4 if seed % 4 == 3 else 1, hasDepth=True, clearColor=(0.0, 0.0, 0.0, 1.0))


P4C4S1 = [(0, 4, 0), (1, 4, 4), (2, 1, 8)]


def points(seed: int) -> Scene:
    """Point lists (DrawCall::setupPoint, Renderer.cpp:1137-1185): gl_PointSize from an attribute, clamped to [1, 1023], squares that
    reach over the frustum sides, perspective w, overlapping points with a depth test or blending, 1x and 4x."""
    rng = np.random.default_rng(23000 + seed)
    n = 40
    w = rng.uniform(0.5, 2.0, n) if seed % 2 else np.ones(n)
    v = np.zeros((n, 9), dtype=np.float32)
    v[:, 0] = rng.uniform(-1.05, 1.05, n) * w
    v[:, 1] = rng.uniform(-1.05, 1.05, n) * w
    v[:, 2] = rng.uniform(0.1, 0.9, n) * w
    v[:, 3] = w
    v[:, 4:8] = rng.uniform(0, 1, (n, 4))
    v[:, 8] = rng.choice([0.25, 1.0, 1.5, 2.0, 3.0, 4.5, 7.0, 12.0, 30.0], n)
    if seed % 6 == 3:
        v[5, 8] = 2000.0  # beyond MAX_POINT_SIZE
    idx = rng.integers(0, n, 25).astype(np.uint16) if seed % 3 == 2 else None
    d = Draw(v, P4C4S1, "vs_point_pos4_col4", "fs_col4", indices=idx, topology=TOPO_POINT_LIST, depthTest=(seed % 2 == 0), depthWrite=True,
             blend=(seed % 4 == 1))
    return Scene(2 * CELL, 2 * CELL, [d], samples=4 if seed % 4 == 2 else 1, hasDepth=True, clearColor=(0.0, 0.0, 0.0, 1.0))


def msaafmt(seed: int) -> Scene:
    """4x MSAA on the colour formats whose resolve is NOT Blitter::fastResolve but the generic blit (Blitter.cpp:2053-2071, :1524-1545):
    sRGB8 (decode per sample, average in linear light, encode), R16G16B16A16_SFLOAT, R32G32B32A32_SFLOAT — opaque, blended and
    depth-tested layers, colours outside [0, 1] on the float targets."""
    rng = np.random.default_rng(25000 + seed)
    fmt = [FMT_R8G8B8A8_SRGB, FMT_R16G16B16A16_SFLOAT, FMT_R32G32B32A32_SFLOAT, FMT_B8G8R8A8_SRGB][seed % 4]
    fl = fmt in (FMT_R16G16B16A16_SFLOAT, FMT_R32G32B32A32_SFLOAT)
    clear = (0.25, 0.5, 0.125, 1.0) if fl else [(0.0, 0.0, 0.0, 1.0), (1.0, 1.0, 1.0, 0.0)][(seed // 4) % 2]
    lo, hi = (-0.75, 2.5) if fl else (0.0, 1.0)
    tris = [_verts(rng, _tri_kind(rng, (5, 0, 1, 4)[i % 4]), persp=(i % 2 == 0), colour=rng.uniform(lo, hi, (3, 4))) for i in range(10)]
    mode = (seed // 4) % 3
    d = Draw(np.concatenate(tris), P4C4, "vs_pos4_col4", "fs_col4", blend=(mode == 1), depthTest=(mode == 2), depthWrite=(mode == 2))
    return Scene(CELL, CELL, [d], samples=4, colorFormat=fmt, clearColor=clear, hasDepth=(mode == 2))


def instanced(seed: int) -> Scene:
    """Instanced draws (CmdDrawBase::draw, VkCommandBuffer.cpp:987-1010: one sw::Renderer::draw per instance, the instance-rate
    streams moved on by Inputs::advanceInstanceAttributes in between): per-instance colour and offset from vertex binding 1."""
    rng = np.random.default_rng(26000 + seed)
    n_inst = 3 + seed % 6
    tris = np.concatenate([_verts(rng, _tri_kind(rng, (1, 5, 4)[i % 3]) * 0.5, persp=(i % 2 == 0)) for i in range(6)])[:, :4].copy()
    inst = np.zeros((n_inst, 8), dtype=np.float32)
    inst[:, 0:4] = rng.uniform(0, 1, (n_inst, 4))
    inst[:, 4:6] = rng.uniform(-0.5, 0.5, (n_inst, 2))
    inst[:, 6] = rng.uniform(-0.1, 0.1, n_inst)
    idx = rng.integers(0, len(tris), 12).astype(np.uint16) if seed % 3 == 1 else None
    d = Draw(tris, [(0, 4, 0)], "vs_inst_pos4_col4", "fs_col4", indices=idx, instances=inst, instanceAttribs=[(1, 4, 0), (2, 4, 4)],
             depthTest=(seed % 2 == 0), depthWrite=True, blend=(seed % 2 == 1))
    return Scene(CELL, CELL, [d], samples=4 if seed % 4 == 3 else 1, hasDepth=True, clearColor=(0.0, 0.0, 0.0, 1.0))


FAMILIES = {
    # name: (generator, number of seeds)
    "coverage": (coverage, 40),
    "zclip": (zclip, 6),
    "cull": (cull, 8),
    "depth_ops": (depth_ops, 16),
    "blend": (blend, 36),
    "stencil": (stencil, 24),
    "texture": (texture, 30),
    "msaa": (msaa, 16),
    "topology": (topology, 12),
    "scissor": (scissor, 8),
    "overdraw": (overdraw, 7),
    "depth16": (depth16, 12),
    "srgb": (srgb, 14),
    "floatrt": (floatrt, 20),
    "pathological": (pathological, 20),
    "srgbtex": (srgbtex, 12),
    "fragtests": (fragtests, 16),
    "texsplit": (texsplit, 12),
    "mixed": (mixed, 24),
    "blendoff": (blendoff, 6),
    "mvp": (mvp, 12),
    "lines": (lines, 16),
    "points": (points, 12),
    "zclamp": (zclamp, 8),
    "msaafmt": (msaafmt, 12),
    "instanced": (instanced, 8),
    "ubo": (ubo, 10),
}


def all_cases():
    for fam, (gen, n) in FAMILIES.items():
        for s in range(n):
            yield f"{fam}_{s}", gen(s)
    for w in range(3):
        yield f"benchmark_{w}_720p", benchmark(w)
    yield "benchmark_0_720p_msaa", benchmark(0, samples=4)
    yield "benchmark_2_720p_msaa", benchmark(2, samples=4)
    # reduced-size instances of the five BASELINE.json workloads (same generators and state as bench.py runs at full size)
    from swiftshader_b200 import workloads
    for w in ("c1", "c2", "c3", "c4", "c5"):
        yield f"workload_{w}", workloads.small(w).scene


def outputs(scene: Scene, att: dict, resolved=None) -> dict:
    """Canonical output arrays for comparison with the reference: unpadded colour (resolved if MSAA), depth, stencil."""
    H = scene.height
    out = {}
    if scene.samples > 1:
        out["color"] = np.ascontiguousarray(resolved[:H])
    else:
        out["color"] = np.ascontiguousarray(att["color"][0, :H])
        if "depth" in att:
            out["depth"] = np.ascontiguousarray(att["depth"][0, :H])
        if "stencil" in att:
            out["stencil"] = np.ascontiguousarray(att["stencil"][0, :H])
    return out
