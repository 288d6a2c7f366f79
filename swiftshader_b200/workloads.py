"""Synthetic workloads of BASELINE.json's configs, generated exactly as SURVEY.md §8(d) specifies (seeded, deterministic).

C1-C3 are the reference's own triangles (/root/reference/tests/VulkanBenchmarks/TriangleBenchmarks.cpp:48-52,100-104,
159-163) at the BASELINE sizes; C4/C5 are the regular-grid meshes.  Each workload carries its algorithmic bytes (the
compulsory DRAM traffic of §8d) so bench.py and DESIGN.md quote the same figure.  Nothing here computes pixels.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .scene import (BF_ONE, BF_ONE_MINUS_SRC_ALPHA, BF_SRC_ALPHA, BF_ZERO, BOP_ADD, CMP_LESS_OR_EQUAL, Draw, Scene, Texture)


def lcg_stream(n: int, seed: int = 0x5EED) -> np.ndarray:
    """First n outputs of x <- 1664525*x + 1013904223 (mod 2^32), vectorised by composing the affine map with itself."""
    out = np.empty(n, dtype=np.uint64)
    a, c = np.uint64(1664525), np.uint64(1013904223)
    m = np.uint64(0xFFFFFFFF)
    out[0] = (a * np.uint64(seed) + c) & m
    k, A, C = 1, a, c  # (A, C): k steps at once
    while k < n:
        cnt = min(k, n - k)
        out[k:k + cnt] = (A * out[:cnt] + C) & m
        C = (A * C + C) & m
        A = (A * A) & m
        k *= 2
    return out.astype(np.uint32)


def _unit(u32: np.ndarray) -> np.ndarray:
    """(x >> 8) / 2^24 in [0, 1)."""
    return (u32 >> 8).astype(np.float64) / float(1 << 24)


@dataclass
class Workload:
    name: str
    scene: Scene
    covered_pixels: int      # pixels shaded per frame (not samples)
    triangles: int
    algorithmic_bytes: int   # whole frame, SURVEY §8d
    tile_bytes: int          # the part moved by the tile kernel (framebuffer read-modify-write)
    description: str


def _grid_indices(nx: int, ny: int) -> np.ndarray:
    j, i = np.meshgrid(np.arange(ny, dtype=np.uint32), np.arange(nx, dtype=np.uint32), indexing="ij")
    v00 = j * (nx + 1) + i
    v10, v01, v11 = v00 + 1, v00 + (nx + 1), v00 + (nx + 2)
    tris = np.stack([v00, v01, v10, v10, v01, v11], axis=-1)  # two CCW triangles per cell (y down)
    return np.ascontiguousarray(tris.reshape(-1), dtype=np.uint32)


def _grid_positions(nx: int, ny: int, width: int, height: int, rnd: np.ndarray) -> np.ndarray:
    """(ny+1)*(nx+1) x 2 NDC positions of a regular grid over the whole viewport, interior vertices jittered by
    <= 0.25 px (so edges are not axis-aligned); border vertices stay on the viewport edge."""
    j, i = np.meshgrid(np.arange(ny + 1), np.arange(nx + 1), indexing="ij")
    x = -1.0 + 2.0 * i / nx
    y = -1.0 + 2.0 * j / ny
    n = (nx + 1) * (ny + 1)
    jx = (_unit(rnd[:n]).reshape(ny + 1, nx + 1) - 0.5) * 2.0 * 0.25 * (2.0 / width)
    jy = (_unit(rnd[n:2 * n]).reshape(ny + 1, nx + 1) - 0.5) * 2.0 * 0.25 * (2.0 / height)
    interior = (i > 0) & (i < nx) & (j > 0) & (j < ny)
    x = x + np.where(interior, jx, 0.0)
    y = y + np.where(interior, jy, 0.0)
    return np.stack([x.reshape(-1), y.reshape(-1)], axis=-1)


def c1(width=1920, height=1080) -> Workload:
    v = np.array([[1, 1, .5], [-1, 1, .5], [0, -1, .5]], dtype=np.float32)
    sc = Scene(width, height, [Draw(v, [(0, 3, 0)], "vs_pos3", "fs_white")], clearColor=(0.5, 0.5, 0.5, 1.0))
    px = width * height // 2
    return Workload("c1_solid_1080p", sc, px, 1, px * 4, px * 4, f"TriangleSolidColor {width}x{height} RGBA8")


def c2(width=1920, height=1080) -> Workload:
    v = np.array([[1, 1, .05, 1, 0, 0], [-1, 1, .5, 0, 1, 0], [0, -1, .5, 0, 0, 1]], dtype=np.float32)
    d = Draw(v, [(0, 3, 0), (1, 3, 3)], "vs_pos3_col3", "fs_col3", depthTest=True, depthWrite=True, depthCompareOp=CMP_LESS_OR_EQUAL)
    sc = Scene(width, height, [d], hasDepth=True, clearDepth=1.0, clearColor=(0.5, 0.5, 0.5, 1.0))
    px = width * height // 2
    return Workload("c2_interp_depth_1080p", sc, px, 1, px * 12, px * 12, f"TriangleInterpolateColor {width}x{height} RGBA8 + D32F LESS_OR_EQUAL")


def checker_texture(size=256) -> Texture:
    """The benchmark's RGB checker generator (TriangleBenchmarks.cpp:219-241) extended to size^2, box-filtered chain."""
    rgb = np.array([0xFFFF0000, 0xFF00FF00, 0xFF0000FF], dtype=np.uint32)
    i, j = np.meshgrid(np.arange(size), np.arange(size), indexing="ij")
    on = ((i ^ j) & 1) == 0
    data = np.zeros(size * size, dtype=np.uint32)
    flat_on = on.reshape(-1)  # i-major order = the order the reference's k counter advances
    k = np.cumsum(flat_on) - 1
    vals = rgb[k % 3]
    idx = (i + size * j).reshape(-1)
    data[idx[flat_on]] = vals[flat_on]
    level0 = data.view(np.uint8).reshape(size, size, 4)
    levels = Texture.box_chain(level0)
    return Texture(levels, maxLod=float(len(levels) - 1))


def c3(width=3840, height=2160) -> Workload:
    v = np.array([[1, 1, .5, 1, 0], [-1, 1, .5, 0, 1], [0, -1, .5, 0, 0]], dtype=np.float32)
    tex = checker_texture(256)
    sc = Scene(width, height, [Draw(v, [(0, 3, 0), (1, 2, 3)], "vs_pos3_uv2", "fs_tex_uv2", texture=tex)], clearColor=(0.5, 0.5, 0.5, 1.0))
    px = width * height // 2
    tb = int(tex.packed().nbytes)
    return Workload("c3_texture_4k", sc, px, 1, px * 4 + tb, px * 4, f"TriangleSampleTexture {width}x{height}, 256^2 RGBA8 9-level trilinear")


def c4(width=3840, height=2160, nx=1000, ny=500) -> Workload:
    """1M small triangles, 4x MSAA, depth test + SRC_ALPHA blending."""
    nv = (nx + 1) * (ny + 1)
    rnd = lcg_stream(nv * 6)
    pos = _grid_positions(nx, ny, width, height, rnd)
    verts = np.zeros((nv, 7), dtype=np.float32)
    verts[:, 0:2] = pos
    verts[:, 2] = 0.25 + 0.5 * _unit(rnd[2 * nv:3 * nv])
    verts[:, 3] = _unit(rnd[3 * nv:4 * nv])
    verts[:, 4] = _unit(rnd[4 * nv:5 * nv])
    verts[:, 5] = _unit(rnd[5 * nv:6 * nv])
    verts[:, 6] = 0.5
    idx = _grid_indices(nx, ny)
    d = Draw(verts, [(0, 3, 0), (1, 4, 3)], "vs_pos3_col4", "fs_col4", indices=idx, depthTest=True, depthWrite=True,
             depthCompareOp=CMP_LESS_OR_EQUAL, blend=True, srcColor=BF_SRC_ALPHA, dstColor=BF_ONE_MINUS_SRC_ALPHA, colorOp=BOP_ADD,
             srcAlpha=BF_ONE, dstAlpha=BF_ZERO, alphaOp=BOP_ADD)
    sc = Scene(width, height, [d], samples=4, hasDepth=True, clearDepth=1.0, clearColor=(0.5, 0.5, 0.5, 1.0))
    px = width * height
    samples = px * 4
    tile_b = samples * 16
    resolve_b = samples * 4 + px * 4
    total = tile_b + verts.nbytes + idx.nbytes + resolve_b
    return Workload("c4_mesh1m_msaa4_blend_4k", sc, px, nx * ny * 2, int(total), int(tile_b),
                    f"{nx * ny * 2} tris ({nx}x{ny} grid), {width}x{height} 4xMSAA RGBA8+D32F, LESS_OR_EQUAL + SRC_ALPHA blend, resolve")


def noise_texture(size=2048, seed=0xC5) -> Texture:
    rng = np.random.default_rng(seed)
    level0 = rng.integers(0, 256, (size, size, 4), dtype=np.uint8)
    levels = Texture.box_chain(level0)
    return Texture(levels, maxLod=float(len(levels) - 1))


def c5(width=7680, height=4320, nx=2500, ny=2000, tex_size=2048) -> Workload:
    """10M-triangle textured mesh at 8K (the multi-GPU band workload)."""
    nv = (nx + 1) * (ny + 1)
    rnd = lcg_stream(nv * 2, seed=0x5EED + 5)
    pos = _grid_positions(nx, ny, width, height, rnd)
    j, i = np.meshgrid(np.arange(ny + 1), np.arange(nx + 1), indexing="ij")
    verts = np.zeros((nv, 5), dtype=np.float32)
    verts[:, 0:2] = pos
    verts[:, 2] = 0.5
    verts[:, 3] = (8.0 * i / nx).reshape(-1)
    verts[:, 4] = (8.0 * j / ny).reshape(-1)
    idx = _grid_indices(nx, ny)
    tex = noise_texture(tex_size)
    d = Draw(verts, [(0, 3, 0), (1, 2, 3)], "vs_pos3_uv2", "fs_tex_uv2", indices=idx, texture=tex)
    sc = Scene(width, height, [d], clearColor=(0.5, 0.5, 0.5, 1.0))
    px = width * height
    total = px * 4 + verts.nbytes + idx.nbytes + tex.packed().nbytes
    return Workload("c5_mesh10m_textured_8k", sc, px, nx * ny * 2, int(total), int(px * 4),
                    f"{nx * ny * 2} tris ({nx}x{ny} grid), {width}x{height} RGBA8, {tex_size}^2 RGBA8 trilinear REPEAT")


WORKLOADS = {"c1": c1, "c2": c2, "c3": c3, "c4": c4, "c5": c5}


def small(name: str) -> Workload:
    """Reduced-size instance of a workload for parity tests (same generator, same state)."""
    if name == "c4":
        return c4(384, 216, 100, 50)
    if name == "c5":
        return c5(512, 288, 160, 128, 128)
    if name == "c3":
        return c3(480, 270)
    if name == "c2":
        return c2(480, 270)
    return c1(480, 270)
