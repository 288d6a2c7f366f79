"""Scenes: the inputs of a draw (vertex/index/texture bytes + pipeline state) and their flattening into
``swcu_draw_desc`` — the Python-side counterpart of what ``sw::Renderer::draw`` gathers
(/root/reference/src/Device/Renderer.cpp:183-490) and of the reference test harness
(tests/VulkanWrapper/DrawTester.cpp:62-102,205-406).

Nothing here computes pixels.  ``Scene.write_ref_scene`` serialises the same inputs for oracle/refrender.cpp
(the harness that drives the reference ICD); ``Device.render`` sends them through the C-ABI to the CUDA path.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import struct
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import capi, spirv

# ---- Vulkan enum values (vulkan_core.h) ----
FMT_R8G8B8A8_UNORM = 37
FMT_B8G8R8A8_UNORM = 44
FMT_R32_SFLOAT, FMT_R32G32_SFLOAT, FMT_R32G32B32_SFLOAT, FMT_R32G32B32A32_SFLOAT = 100, 103, 106, 109
FMT_R16G16B16A16_SFLOAT = 97
FMT_R8G8B8A8_SRGB = 43
FMT_B8G8R8A8_SRGB = 50
FMT_D32_SFLOAT = 126
FMT_D16_UNORM = 124
FMT_S8_UINT = 127
FLOAT_FORMATS = {1: FMT_R32_SFLOAT, 2: FMT_R32G32_SFLOAT, 3: FMT_R32G32B32_SFLOAT, 4: FMT_R32G32B32A32_SFLOAT}
TOPO_POINT_LIST, TOPO_LINE_LIST, TOPO_LINE_STRIP = 0, 1, 2
TOPO_TRIANGLE_LIST, TOPO_TRIANGLE_STRIP, TOPO_TRIANGLE_FAN = 3, 4, 5
CMP_NEVER, CMP_LESS, CMP_EQUAL, CMP_LESS_OR_EQUAL, CMP_GREATER, CMP_NOT_EQUAL, CMP_GREATER_OR_EQUAL, CMP_ALWAYS = range(8)
SOP_KEEP, SOP_ZERO, SOP_REPLACE, SOP_INC_CLAMP, SOP_DEC_CLAMP, SOP_INVERT, SOP_INC_WRAP, SOP_DEC_WRAP = range(8)
(BF_ZERO, BF_ONE, BF_SRC_COLOR, BF_ONE_MINUS_SRC_COLOR, BF_DST_COLOR, BF_ONE_MINUS_DST_COLOR, BF_SRC_ALPHA,
 BF_ONE_MINUS_SRC_ALPHA, BF_DST_ALPHA, BF_ONE_MINUS_DST_ALPHA, BF_CONSTANT_COLOR, BF_ONE_MINUS_CONSTANT_COLOR,
 BF_CONSTANT_ALPHA, BF_ONE_MINUS_CONSTANT_ALPHA, BF_SRC_ALPHA_SATURATE) = range(15)
BOP_ADD, BOP_SUBTRACT, BOP_REVERSE_SUBTRACT, BOP_MIN, BOP_MAX = range(5)
CULL_NONE, CULL_FRONT, CULL_BACK = 0, 1, 2
FRONT_CCW, FRONT_CW = 0, 1
FILTER_NEAREST, FILTER_LINEAR = 0, 1
MIPMAP_NEAREST, MIPMAP_LINEAR = 0, 1
ADDR_REPEAT, ADDR_MIRRORED_REPEAT, ADDR_CLAMP_TO_EDGE = 0, 1, 2


@dataclass
class StencilFace:
    failOp: int = SOP_KEEP
    passOp: int = SOP_KEEP
    depthFailOp: int = SOP_KEEP
    compareOp: int = CMP_ALWAYS
    compareMask: int = 0xFF
    writeMask: int = 0xFF
    reference: int = 0

    def tuple(self):
        return (self.failOp, self.passOp, self.depthFailOp, self.compareOp, self.compareMask, self.writeMask, self.reference)


@dataclass
class Texture:
    """RGBA8 2-D texture with a mip chain + sampler state (vk::SamplerState, src/Vulkan/VkSampler.hpp:29-63)."""
    levels: list  # list of (h, w, 4) uint8 arrays, level 0 first
    magFilter: int = FILTER_LINEAR
    minFilter: int = FILTER_LINEAR
    mipmapMode: int = MIPMAP_LINEAR
    addressModeU: int = ADDR_REPEAT
    addressModeV: int = ADDR_REPEAT
    mipLodBias: float = 0.0
    minLod: float = 0.0
    maxLod: float = 0.0
    set: int = 0
    binding: int = 0
    srgb: bool = False  # R8G8B8A8_SRGB instead of R8G8B8A8_UNORM

    def packed(self) -> np.ndarray:
        """All levels back to back, tightly packed (the layout both the ICD upload and our desc use)."""
        return np.concatenate([np.ascontiguousarray(l, dtype=np.uint8).reshape(-1) for l in self.levels])

    @staticmethod
    def box_chain(level0: np.ndarray) -> list:
        """Full mip chain by 2x2 box filtering (host-side, like an application would upload)."""
        levels = [np.ascontiguousarray(level0, dtype=np.uint8)]
        while levels[-1].shape[0] > 1 or levels[-1].shape[1] > 1:
            a = levels[-1].astype(np.uint32)
            h, w = a.shape[:2]
            h2, w2 = max(1, h // 2), max(1, w // 2)
            a = a[: h2 * 2 if h > 1 else 1, : w2 * 2 if w > 1 else 1]
            if h > 1:
                a = a[0::2] + a[1::2]
            else:
                a = a * 2
            if w > 1:
                a = a[:, 0::2] + a[:, 1::2]
            else:
                a = a * 2
            levels.append(((a + 2) // 4).astype(np.uint8))
        return levels


@dataclass
class Draw:
    vertices: np.ndarray  # (N, k) float32, interleaved
    attribs: list  # [(location, components, float_offset)]
    vs: str
    fs: str
    indices: Optional[np.ndarray] = None  # uint16 / uint32, or None
    topology: int = TOPO_TRIANGLE_LIST
    count: Optional[int] = None  # vertex/index count (default: all)
    first: int = 0  # firstIndex / firstVertex
    vertexOffset: int = 0
    viewport: Optional[tuple] = None  # (x, y, w, h, minDepth, maxDepth); default full FB
    scissor: Optional[tuple] = None  # (x, y, w, h)
    cullMode: int = CULL_NONE
    frontFace: int = FRONT_CCW
    depthTest: bool = False
    depthWrite: bool = False
    depthCompareOp: int = CMP_LESS_OR_EQUAL
    stencilTest: bool = False
    front: StencilFace = field(default_factory=StencilFace)
    back: StencilFace = field(default_factory=StencilFace)
    depthBias: tuple = (0.0, 0.0, 0.0)  # constant, clamp, slope
    blend: bool = False
    srcColor: int = BF_SRC_ALPHA
    dstColor: int = BF_ONE_MINUS_SRC_ALPHA
    colorOp: int = BOP_ADD
    srcAlpha: int = BF_ONE
    dstAlpha: int = BF_ZERO
    alphaOp: int = BOP_ADD
    colorWriteMask: int = 0xF
    blendConstants: tuple = (0.0, 0.0, 0.0, 0.0)
    sampleMask: int = 0xFFFFFFFF
    alphaToCoverage: bool = False
    depthBounds: Optional[tuple] = None  # (min, max) enables the depth bounds test
    texture: Optional[Texture] = None
    pushConstants: Optional[np.ndarray] = None  # float32 / uint32 words pushed for the vertex stage (<= 32)
    # uniform buffer the vertex stage reads: (set, binding, float32 / uint32 words) — VK_DESCRIPTOR_TYPE_UNIFORM_BUFFER
    uniformBuffer: Optional[tuple] = None
    lineWidth: float = 1.0
    depthClamp: bool = False  # depthClampEnable: no near / far clipping, fragment depth clamped to the viewport's depth range
    # instancing (CmdDrawBase::draw, VkCommandBuffer.cpp:987-1010): one Renderer::draw per instance, the streams of the instance-rate
    # binding moved on by its stride in between (Inputs::advanceInstanceAttributes, Context.cpp:447-461); the vertices see stride 0
    instances: Optional[np.ndarray] = None  # (instanceCount, k) float32: the instance-rate vertex buffer (binding 1)
    instanceAttribs: list = field(default_factory=list)  # [(location, components, float_offset)] inside a row of `instances`
    instanceRow: Optional[int] = None  # set on the per-instance copies Scene.flat_draws() hands out

    def vertex_count(self) -> int:
        if self.count is not None:
            return self.count
        return len(self.indices) if self.indices is not None else self.vertices.shape[0]

    def primitive_count(self) -> int:
        n = self.vertex_count()
        if self.topology == TOPO_TRIANGLE_LIST:
            return n // 3
        if self.topology == TOPO_POINT_LIST:
            return n
        if self.topology == TOPO_LINE_LIST:
            return n // 2
        if self.topology == TOPO_LINE_STRIP:
            return max(0, n - 1)
        return max(0, n - 2)


@dataclass
class Scene:
    width: int
    height: int
    draws: list
    samples: int = 1
    colorFormat: int = FMT_R8G8B8A8_UNORM
    hasDepth: bool = False
    hasStencil: bool = False
    depthFormat: int = FMT_D32_SFLOAT  # or FMT_D16_UNORM (then no stencil)
    clearColor: tuple = (0.0, 0.0, 0.0, 0.0)
    clearDepth: float = 1.0
    clearStencil: int = 0

    # ---- attachments (linear rows, height padded to even, sample q = slice q) ----
    def padded_height(self) -> int:
        return (self.height + 1) & ~1

    def color_dtype(self) -> np.dtype:
        """Element type of one colour channel: uint8 (RGBA8 / BGRA8, UNORM or SRGB), float16 (R16G16B16A16_SFLOAT) or float32
        (R32G32B32A32_SFLOAT); a pixel is four of them."""
        return np.dtype({FMT_R32G32B32A32_SFLOAT: np.float32, FMT_R16G16B16A16_SFLOAT: np.float16}.get(self.colorFormat, np.uint8))

    def color_bpp(self) -> int:
        return 4 * self.color_dtype().itemsize

    def clear_color_bytes(self) -> np.ndarray:
        """One pixel of the clear colour in the attachment's format.  UNORM8: the pack Blitter::fastClear does,
        (uint32_t)(255 * c + 0.5f) (/root/reference/src/Device/Blitter.cpp:217-229); float formats: the value itself (use clear
        colours that are exact in half precision for R16G16B16A16_SFLOAT)."""
        if self.color_dtype() != np.uint8:
            return np.array(self.clearColor, dtype=np.float32).astype(self.color_dtype())
        c = np.clip(np.array(self.clearColor, dtype=np.float32), 0, 1)
        b = (np.float32(255.0) * c + np.float32(0.5)).astype(np.float32).astype(np.uint32).astype(np.uint8)
        if self.colorFormat in (FMT_B8G8R8A8_UNORM, FMT_B8G8R8A8_SRGB):
            b = b[[2, 1, 0, 3]]  # (sRGB targets: use clear colours of 0 / 1 only — the clear's own sRGB encode is not restated here)
        return b

    def alloc_attachments(self) -> dict:
        H2, W, S = self.padded_height(), self.width, self.samples
        att = {"color": np.empty((S, H2, W, 4), dtype=self.color_dtype())}
        att["color"][:] = self.clear_color_bytes()
        if self.hasDepth and self.depthFormat == FMT_D16_UNORM:
            # D16 is not a fastClear format: the generic blit scales by 0xFFFF, clamps and stores UShort(RoundInt(x))
            # (/root/reference/src/Device/Blitter.cpp:1085-1087), i.e. round-to-nearest-even
            z16 = np.uint16(np.rint(np.float32(65535.0) * np.float32(min(max(self.clearDepth, 0.0), 1.0))))
            att["depth"] = np.full((S, H2, W), z16, dtype=np.uint16)
        elif self.hasDepth:
            att["depth"] = np.full((S, H2, W), self.clearDepth, dtype=np.float32)
        if self.hasStencil:
            att["stencil"] = np.full((S, H2, W), self.clearStencil & 0xFF, dtype=np.uint8)
        return att

    def flat_draws(self) -> list:
        """The draws as sw::Renderer::draw sees them: an instanced draw is one draw per instance."""
        out = []
        for dr in self.draws:
            if dr.instances is None:
                out.append(dr)
            else:
                out += [dataclasses.replace(dr, instanceRow=i) for i in range(dr.instances.shape[0])]
        return out

    # ---- flattening into the C-ABI descriptor ----
    def build_desc(self, draw: Draw, att: dict, keep: list, render_area: Optional[tuple] = None,
                   dev: Optional[list] = None) -> capi.DrawDesc:
        """Fill a swcu_draw_desc with HOST addresses of numpy buffers. ``keep`` receives every array that must
        stay alive while the descriptor is in use; ``dev`` the subset the device reads (needs a shadow)."""
        dev = dev if dev is not None else []
        d = capi.DrawDesc()
        d.structSize = C.sizeof(capi.DrawDesc)
        d.topology = draw.topology
        d.provokingVertexMode = 0
        verts = np.ascontiguousarray(draw.vertices, dtype=np.float32)
        keep.append(verts)
        dev.append(verts)
        stride = verts.shape[1] * 4
        nverts = draw.vertex_count()
        if draw.indices is not None:
            idx = np.ascontiguousarray(draw.indices)
            assert idx.dtype in (np.uint16, np.uint32)
            keep.append(idx)
            dev.append(idx)
            d.indexType = idx.dtype.itemsize
            d.indexBuffer = idx.ctypes.data + draw.first * idx.dtype.itemsize
            d.baseVertex = draw.vertexOffset
        else:
            d.indexType = 0
            d.indexBuffer = None
            d.baseVertex = draw.first
        d.primitiveCount = draw.primitive_count()
        d.lineWidth = draw.lineWidth
        if draw.pushConstants is not None:
            pc = np.ascontiguousarray(draw.pushConstants).view(np.uint32).ravel().copy()
            keep.append(pc)
            d.pushConstants, d.pushConstantBytes = pc.ctypes.data, pc.nbytes
        if draw.uniformBuffer is not None:
            (uset, ubinding, uwords) = draw.uniformBuffer
            ub = np.ascontiguousarray(uwords).view(np.uint32).ravel().copy()
            keep.append(ub)  # host memory: swcu_draw folds the words the vertex program reads (no shadow)
            d.uniformBufferCount = 1
            d.uniformBuffer[0] = capi.UniformBuffer(uset, ubinding, ub.ctypes.data, ub.nbytes, 0)
        for (loc, comps, off) in draw.attribs:
            vi = d.input[loc]
            vi.buffer = verts.ctypes.data + off * 4
            vi.robustnessSize = max(0, verts.nbytes - off * 4)
            vi.vertexStride = stride
            vi.format = FLOAT_FORMATS[comps]
        if draw.instances is not None:
            inst = np.ascontiguousarray(draw.instances, dtype=np.float32)
            keep.append(inst)
            dev.append(inst)
            row = draw.instanceRow or 0
            for (loc, comps, off) in draw.instanceAttribs:
                vi = d.input[loc]
                start = (row * inst.shape[1] + off) * 4
                vi.buffer = inst.ctypes.data + start
                vi.robustnessSize = max(0, inst.nbytes - start)
                vi.vertexStride = 0
                vi.format = FLOAT_FORMATS[comps]
        vs, fs = spirv.shader(draw.vs), spirv.shader(draw.fs)
        keep += [vs, fs]
        d.vertexShader, d.vertexShaderWords = vs.ctypes.data, len(vs)
        d.fragmentShader, d.fragmentShaderWords = fs.ctypes.data, len(fs)
        vp = draw.viewport or (0.0, 0.0, float(self.width), float(self.height), 0.0, 1.0)
        (d.viewportX, d.viewportY, d.viewportWidth, d.viewportHeight, d.viewportMinDepth, d.viewportMaxDepth) = vp
        sc = draw.scissor or (0, 0, self.width, self.height)
        d.scissor = capi.Rect(*sc)
        d.renderArea = capi.Rect(*(render_area or (0, 0, self.width, self.height)))
        d.cullMode, d.frontFace = draw.cullMode, draw.frontFace
        d.depthClampEnable, d.depthClipEnable = int(draw.depthClamp), int(not draw.depthClamp)  # Context.cpp:647-648
        d.depthBiasConstant, d.depthBiasClamp, d.depthBiasSlope = draw.depthBias
        d.sampleCount, d.sampleMask = self.samples, draw.sampleMask & ((1 << self.samples) - 1)
        d.alphaToCoverageEnable = int(draw.alphaToCoverage)
        if draw.depthBounds is not None:
            d.depthBoundsTestEnable, (d.minDepthBounds, d.maxDepthBounds) = 1, draw.depthBounds
        d.depthTestEnable, d.depthWriteEnable, d.depthCompareOp = int(draw.depthTest), int(draw.depthWrite), draw.depthCompareOp
        d.stencilTestEnable = int(draw.stencilTest)
        d.front = capi.StencilFace(*draw.front.tuple())
        d.back = capi.StencilFace(*draw.back.tuple())
        d.blendEnable = int(draw.blend)
        d.srcColorBlendFactor, d.dstColorBlendFactor, d.colorBlendOp = draw.srcColor, draw.dstColor, draw.colorOp
        d.srcAlphaBlendFactor, d.dstAlphaBlendFactor, d.alphaBlendOp = draw.srcAlpha, draw.dstAlpha, draw.alphaOp
        d.colorWriteMask = draw.colorWriteMask
        d.blendConstants = (C.c_float * 4)(*draw.blendConstants)
        H2, W = self.padded_height(), self.width
        col = att["color"]
        cb = self.color_bpp()
        d.color = capi.Attachment(col.ctypes.data, self.colorFormat, W * cb, H2 * W * cb, W, self.height, 0)
        if "depth" in att:
            zb = att["depth"].dtype.itemsize
            d.depth = capi.Attachment(att["depth"].ctypes.data, self.depthFormat, W * zb, H2 * W * zb, W, self.height, 0)
        if "stencil" in att:
            d.stencil = capi.Attachment(att["stencil"].ctypes.data, FMT_S8_UINT, W, H2 * W, W, self.height, 0)
        if draw.texture is not None:
            t = draw.texture
            packed = t.packed()
            keep.append(packed)
            dev.append(packed)
            si = d.sampledImage[0]
            si.set, si.binding, si.format, si.levelCount = t.set, t.binding, (FMT_R8G8B8A8_SRGB if t.srgb else FMT_R8G8B8A8_UNORM), len(t.levels)
            off = 0
            for l in range(capi.MIPMAP_LEVELS):
                ll = min(l, len(t.levels) - 1)
                if l < len(t.levels):
                    cur = off
                    off += t.levels[l].shape[0] * t.levels[l].shape[1] * 4
                    last = cur
                h, w = t.levels[ll].shape[:2]
                si.level[l] = capi.MipLevel(packed.ctypes.data + (cur if l < len(t.levels) else last), w, h, w, 0)
            si.magFilter, si.minFilter, si.mipmapMode = t.magFilter, t.minFilter, t.mipmapMode
            si.addressModeU, si.addressModeV = t.addressModeU, t.addressModeV
            si.mipLodBias, si.minLod, si.maxLod = t.mipLodBias, t.minLod, t.maxLod
            d.sampledImageCount = 1
        return d

    # ---- serialisation for oracle/refrender.cpp (oracle/scene_format.h) ----
    def write_ref_scene(self, path: str) -> None:
        blobs: list = []

        def blob(a) -> int:
            blobs.append(np.ascontiguousarray(a).view(np.uint8).reshape(-1))
            return len(blobs) - 1

        recs = []
        for dr in self.draws:
            verts = np.ascontiguousarray(dr.vertices, dtype=np.float32)
            r = struct.pack("<IIIIi", dr.topology, 0 if dr.indices is None else dr.indices.dtype.itemsize,
                            dr.vertex_count(), dr.first, dr.vertexOffset)
            r += struct.pack("<IIII", blob(spirv.shader(dr.vs)), blob(spirv.shader(dr.fs)), blob(verts),
                             blob(dr.indices) if dr.indices is not None else 0)
            r += struct.pack("<II", verts.shape[1] * 4, len(dr.attribs))
            for i in range(8):
                if i < len(dr.attribs):
                    loc, comps, off = dr.attribs[i]
                    r += struct.pack("<III", loc, FLOAT_FORMATS[comps], off * 4)
                else:
                    r += struct.pack("<III", 0, 0, 0)
            vp = dr.viewport or (0.0, 0.0, float(self.width), float(self.height), 0.0, 1.0)
            sc = dr.scissor or (0, 0, self.width, self.height)
            r += struct.pack("<6f4i", *vp, *sc)
            r += struct.pack("<IIIIII", dr.cullMode, dr.frontFace, int(dr.depthTest), int(dr.depthWrite), dr.depthCompareOp,
                             int(dr.stencilTest))
            r += struct.pack("<7I7I", *dr.front.tuple(), *dr.back.tuple())
            r += struct.pack("<Ifff", int(any(v != 0.0 for v in dr.depthBias)), *dr.depthBias)
            r += struct.pack("<8I4fI", int(dr.blend), dr.srcColor, dr.dstColor, dr.colorOp, dr.srcAlpha, dr.dstAlpha, dr.alphaOp,
                             dr.colorWriteMask, *dr.blendConstants, dr.sampleMask & 0xFFFFFFFF)
            t = dr.texture
            if t is not None:
                r += struct.pack("<5I5I3f2I", 2 if t.srgb else 1, blob(t.packed()), t.levels[0].shape[1], t.levels[0].shape[0], len(t.levels),
                                 t.magFilter, t.minFilter, t.mipmapMode, t.addressModeU, t.addressModeV,
                                 t.mipLodBias, t.minLod, t.maxLod, t.set, t.binding)
            else:
                r += struct.pack("<5I5I3f2I", 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0.0, 0.0, 0.0, 0, 0)
            r += struct.pack("<IIff", int(dr.alphaToCoverage), int(dr.depthBounds is not None), *(dr.depthBounds or (0.0, 1.0)))
            pc = np.zeros(32, dtype=np.uint32)
            npc = 0
            if dr.pushConstants is not None:
                w = np.ascontiguousarray(dr.pushConstants).view(np.uint32).ravel()
                npc = len(w)
                pc[:npc] = w
            r += struct.pack("<fI", dr.lineWidth, 4 * npc) + pc.tobytes()
            r += struct.pack("<I", int(dr.depthClamp))
            # instancing: the instance-rate buffer (binding 1) and its attributes
            if dr.instances is not None:
                inst = np.ascontiguousarray(dr.instances, dtype=np.float32)
                r += struct.pack("<IIII", inst.shape[0], blob(inst), inst.shape[1] * 4, len(dr.instanceAttribs))
                for i in range(4):
                    if i < len(dr.instanceAttribs):
                        loc, comps, off = dr.instanceAttribs[i]
                        r += struct.pack("<III", loc, FLOAT_FORMATS[comps], off * 4)
                    else:
                        r += struct.pack("<III", 0, 0, 0)
            else:
                r += struct.pack("<IIII", 1, 0, 0, 0) + struct.pack("<III", 0, 0, 0) * 4
            if dr.uniformBuffer is not None:
                (uset, ubinding, uwords) = dr.uniformBuffer
                r += struct.pack("<IIII", 1, blob(np.ascontiguousarray(uwords).view(np.uint32).ravel()), uset, ubinding)
            else:
                r += struct.pack("<IIII", 0, 0, 0, 0)
            recs.append(r)
        hdr = struct.pack("<IIIIIIII4ffIII", 0x43535753, 7, self.width, self.height, self.samples, self.colorFormat,
                          (2 if self.depthFormat == FMT_D16_UNORM else 1) if self.hasDepth else 0, int(self.hasStencil), *self.clearColor, self.clearDepth, self.clearStencil,
                          len(recs), len(blobs))
        off = len(hdr) + sum(len(r) for r in recs) + 16 * len(blobs)
        table = b""
        for b in blobs:
            off = (off + 15) & ~15
            table += struct.pack("<QQ", off, b.nbytes)
            off += b.nbytes
        with open(path, "wb") as f:
            f.write(hdr)
            for r in recs:
                f.write(r)
            f.write(table)
            pos = len(hdr) + sum(len(r) for r in recs) + len(table)
            for b in blobs:
                pad = ((pos + 15) & ~15) - pos
                f.write(b"\0" * pad)
                pos += pad
                f.write(b.tobytes())
                pos += b.nbytes


def resolve_host(color: np.ndarray) -> np.ndarray:
    """Shape helper only (no pixel math): the 1x view of a 1-sample attachment."""
    return color[0]


class Device:
    """One CUDA context of the C-ABI (``swcu_ctx``): the Python stand-in for the shim a maintainer would put in
    ``DrawCall::run``.  Holds the device shadows of every host buffer it has seen."""

    def __init__(self, ordinal: int = 0):
        self.lib = capi.lib()
        self.ctx = C.c_void_p()
        rc = self.lib.swcu_create(C.byref(self.ctx), ordinal)
        if rc != capi.OK:
            raise capi.SwcuError(rc, (self.lib.swcu_last_error(None) or b"").decode())
        self._registered: dict = {}

    def close(self):
        if self.ctx:
            self.lib.swcu_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc: int):
        if rc != capi.OK:
            raise capi.SwcuError(rc, (self.lib.swcu_last_error(self.ctx) or b"").decode())

    # memory
    def register(self, arr: np.ndarray, upload: bool = True):
        key = arr.ctypes.data
        if key not in self._registered:
            self.check(self.lib.swcu_mem_register(self.ctx, key, arr.nbytes))
            self._registered[key] = arr
        if upload:
            self.check(self.lib.swcu_mem_upload(self.ctx, key, arr.nbytes))

    def unregister(self, arr: np.ndarray):
        key = arr.ctypes.data
        if key in self._registered:
            self.check(self.lib.swcu_mem_unregister(self.ctx, key))
            del self._registered[key]

    def download(self, arr: np.ndarray):
        self.check(self.lib.swcu_mem_download(self.ctx, arr.ctypes.data, arr.nbytes))

    def device_ptr(self, arr: np.ndarray) -> int:
        return self.lib.swcu_mem_device_ptr(self.ctx, arr.ctypes.data) or 0

    def sync(self):
        self.check(self.lib.swcu_sync(self.ctx))

    def fence_signal(self, slot: int):
        """Marks everything issued so far (copies included); ``fence_wait`` blocks the host on it without draining the device."""
        self.check(self.lib.swcu_fence_signal(self.ctx, slot))

    def fence_wait(self, slot: int):
        self.check(self.lib.swcu_fence_wait(self.ctx, slot))

    def draw(self, desc: capi.DrawDesc):
        self.check(self.lib.swcu_draw(self.ctx, C.byref(desc)))

    def upload(self, arr: np.ndarray):
        self.check(self.lib.swcu_mem_upload(self.ctx, arr.ctypes.data, arr.nbytes))

    def set_stream(self, cuda_stream: int):
        self.check(self.lib.swcu_set_stream(self.ctx, cuda_stream))

    def reset_stats(self):
        self.check(self.lib.swcu_reset_stats(self.ctx))

    def set_profiling(self, on):
        self.check(self.lib.swcu_set_profiling(self.ctx, int(on)))

    def last_draw_kernels(self) -> list:
        names = (C.c_char_p * 64)()
        ms = (C.c_float * 64)()
        n = self.lib.swcu_last_draw_kernels(self.ctx, names, ms, 64)
        return [(names[i].decode(), float(ms[i])) for i in range(max(n, 0))]

    def timeline(self, cap: int = 4096) -> list:
        """(name, begin ms, end ms) of every kernel issued since set_profiling(2), on the streams they really ran on."""
        names = (C.c_char_p * cap)()
        t0 = (C.c_float * cap)()
        t1 = (C.c_float * cap)()
        n = self.lib.swcu_timeline(self.ctx, names, t0, t1, cap)
        return [(names[i].decode(), float(t0[i]), float(t1[i])) for i in range(max(n, 0))]

    def set_option(self, name: str, value: int):
        self.check(self.lib.swcu_set_option(self.ctx, name.encode(), value))

    def stats(self) -> capi.Stats:
        s = capi.Stats()
        self.check(self.lib.swcu_get_stats(self.ctx, C.byref(s)))
        return s

    def resolve(self, scene: Scene, att: dict) -> np.ndarray:
        """Blitter::fastResolve on the device shadows; returns the 1x image."""
        H2, W = scene.padded_height(), scene.width
        out = np.zeros((1, H2, W, 4), dtype=scene.color_dtype())
        bpp = scene.color_bpp()
        self.register(out, upload=True)
        src = capi.Attachment(att["color"].ctypes.data, scene.colorFormat, W * bpp, H2 * W * bpp, W, scene.height, 0)
        dst = capi.Attachment(out.ctypes.data, scene.colorFormat, W * bpp, H2 * W * bpp, W, scene.height, 0)
        self.check(self.lib.swcu_resolve(self.ctx, C.byref(src), scene.samples, C.byref(dst)))
        self.download(out)
        self.sync()
        self.unregister(out)
        return out[0]

    def render(self, scene: Scene, att: Optional[dict] = None, render_area: Optional[tuple] = None) -> dict:
        """Upload inputs, issue every draw of the scene through swcu_draw, read the attachments back."""
        att = att if att is not None else scene.alloc_attachments()
        keep: list = []
        dev: list = []
        descs = [scene.build_desc(dr, att, keep, render_area, dev) for dr in scene.flat_draws()]
        bufs, seen = [], set()
        for b in list(att.values()) + dev:
            if b.ctypes.data not in seen:
                seen.add(b.ctypes.data)
                bufs.append(b)
        for b in bufs:
            self.register(b)
        for d in descs:
            self.draw(d)
        for a in att.values():
            self.download(a)
        self.sync()
        for b in bufs:
            self.unregister(b)
        return att


class Frame:
    """One scene kept resident on a Device: buffers are registered once (device shadows allocated, host side page-locked),
    then ``upload_inputs`` / ``clear`` / ``draw`` / ``resolve`` / ``download`` can be issued any number of times.  This is
    the steady-state shape of a render loop behind ``sw::Renderer::draw`` (inputs uploaded at vkQueueSubmit, attachments
    resident across frames)."""

    def __init__(self, dev: Device, scene: Scene, render_area: Optional[tuple] = None):
        self.dev, self.scene = dev, scene
        self.att = scene.alloc_attachments()
        H2, W = scene.padded_height(), scene.width
        self.resolved = np.zeros((1, H2, W, 4), dtype=scene.color_dtype()) if scene.samples > 1 else None
        self.keep: list = []
        self.inputs: list = []
        self.descs = [scene.build_desc(dr, self.att, self.keep, render_area, self.inputs) for dr in scene.flat_draws()]
        seen = set()
        self.inputs = [b for b in self.inputs if not (b.ctypes.data in seen or seen.add(b.ctypes.data))]
        self.bufs = list(self.att.values()) + self.inputs + ([self.resolved] if self.resolved is not None else [])
        for b in self.bufs:
            dev.register(b, upload=False)
        self.render_area = render_area or (0, 0, scene.width, scene.height)

    def input_bytes(self) -> int:
        return int(sum(b.nbytes for b in self.inputs))

    def upload_inputs(self):
        for b in self.inputs:
            self.dev.upload(b)

    def upload_attachments(self):
        for b in self.att.values():
            self.dev.upload(b)

    def _attachment(self, key: str) -> capi.Attachment:
        H2, W, sc = self.scene.padded_height(), self.scene.width, self.scene
        cb = sc.color_bpp()
        if key == "color":
            return capi.Attachment(self.att["color"].ctypes.data, sc.colorFormat, W * cb, H2 * W * cb, W, sc.height, 0)
        if key == "depth":
            zb = self.att["depth"].dtype.itemsize
            return capi.Attachment(self.att["depth"].ctypes.data, sc.depthFormat, W * zb, H2 * W * zb, W, sc.height, 0)
        if key == "stencil":
            return capi.Attachment(self.att["stencil"].ctypes.data, FMT_S8_UINT, W, H2 * W, W, sc.height, 0)
        return capi.Attachment(self.resolved.ctypes.data, sc.colorFormat, W * cb, H2 * W * cb, W, sc.height, 0)

    def clear(self):
        """Attachment load-op CLEAR on the device (Blitter::fastClear), whole framebuffer."""
        sc = self.scene
        area = capi.Rect(0, 0, sc.width, sc.height)
        col = sc.clear_color_bytes().copy()
        a = self._attachment("color")
        self.dev.check(self.dev.lib.swcu_clear(self.dev.ctx, C.byref(a), sc.samples, C.byref(area), col.ctypes.data))
        if "depth" in self.att:
            if sc.depthFormat == FMT_D16_UNORM:  # same value as the host-side clear of alloc_attachments
                z = np.array([np.rint(np.float32(65535.0) * np.float32(min(max(sc.clearDepth, 0.0), 1.0)))], dtype=np.uint16)
            else:
                z = np.array([sc.clearDepth], dtype=np.float32)
            a = self._attachment("depth")
            self.dev.check(self.dev.lib.swcu_clear(self.dev.ctx, C.byref(a), sc.samples, C.byref(area), z.ctypes.data))
        if "stencil" in self.att:
            s = np.array([sc.clearStencil & 0xFF, 0, 0, 0], dtype=np.uint8)
            a = self._attachment("stencil")
            self.dev.check(self.dev.lib.swcu_clear(self.dev.ctx, C.byref(a), sc.samples, C.byref(area), s.ctypes.data))

    def draw(self):
        for d in self.descs:
            self.dev.draw(d)

    def resolve(self):
        if self.resolved is None:
            return
        src, dst = self._attachment("color"), self._attachment("resolved")
        self.dev.check(self.dev.lib.swcu_resolve(self.dev.ctx, C.byref(src), self.scene.samples, C.byref(dst)))

    def final_image(self) -> np.ndarray:
        """Host array that holds the presentable 1x image after download_final()."""
        return self.resolved[0] if self.resolved is not None else self.att["color"][0]

    def download_final(self, rows: Optional[tuple] = None):
        img = self.final_image()
        if rows is None:
            self.dev.download(img)
        else:
            self.dev.download(img[rows[0]:rows[1]])

    def download_all(self):
        for b in self.att.values():
            self.dev.download(b)
        if self.resolved is not None:
            self.dev.download(self.resolved)

    def final_device_ptr(self) -> int:
        return self.dev.device_ptr(self.final_image())

    def close(self):
        self.dev.sync()
        for b in self.bufs:
            self.dev.unregister(b)
        self.bufs = []
