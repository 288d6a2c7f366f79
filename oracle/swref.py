"""Python binding of the CPU oracle (oracle/libswref.so) and of the reference harness (oracle/_ref/refrender).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing under swiftshader_b200/ imports this module.

The oracle does not parse SPIR-V: the meaning of each fixture shader is written down by hand in SHADER_SPECS
(from the GLSL in /root/reference/tests/VulkanBenchmarks/TriangleBenchmarks.cpp:56-81,109-140,168-198 and
SURVEY.md §9.2), so that the product's translator (csrc/spirv_subset.cpp) is checked against an independent
statement of what the shader does.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import struct
import subprocess
import tempfile

import numpy as np

from swiftshader_b200 import capi
from swiftshader_b200.scene import Scene

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libswref.so")
REF_ICD = os.path.join(_HERE, "_ref", "libvk_swiftshader.so")
REFRENDER = os.path.join(_HERE, "_ref", "refrender")
# the same ICD with the CUDA draw path wired in behind sw::Renderer::draw (oracle/build_cuda_icd.sh, icd/): product, not oracle —
# it is only DRIVEN from here, by the harness that also drives the reference
CUDA_ICD = os.path.join(_HERE, "_cuda", "libvk_swiftshader_cuda.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "libswref.so"])


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.swref_draw.argtypes = [C.POINTER(capi.DrawDesc), C.POINTER(capi.ShaderInfo), C.POINTER(capi.ShaderInfo)]
        L.swref_clear.argtypes = [C.POINTER(capi.Attachment), C.c_uint32, C.POINTER(capi.Rect), C.c_void_p]
        L.swref_resolve.argtypes = [C.POINTER(capi.Attachment), C.c_uint32, C.POINTER(capi.Attachment)]
        L.swref_version.restype = C.c_char_p
        _lib = L
    return _lib


def _f32bits(x: float) -> int:
    return struct.unpack("<I", struct.pack("<f", x))[0]


def _inp(loc: int, comp: int):
    return (capi.SRC_INPUT, loc * 4 + comp)


def _const(x: float):
    return (capi.SRC_CONST, _f32bits(x))


def _texel(c: int):
    return (capi.SRC_TEXEL, c)


def _push(word: int):
    return (capi.SRC_PUSH, word)


def _ubo(slot: int, word: int):
    return (capi.SRC_UNIFORM, (slot << 16) | word)


def _temp(step: int):
    return (capi.SRC_TEMP, step)


class _Program:
    """The scalar steps of a vertex shader's arithmetic, written down by hand in the reference's evaluation order."""

    def __init__(self):
        self.steps = []

    def op(self, op, a, b=None, c=None):
        z = _const(0.0)
        self.steps.append((op, a, b or z, c or z))
        return _temp(len(self.steps) - 1)

    def mul(self, a, b):
        return self.op(capi.OP_MUL, a, b)

    def add(self, a, b):
        return self.op(capi.OP_ADD, a, b)

    def fma(self, a, b, c):
        return self.op(capi.OP_FMA, a, b, c)

    def mat4_times_vec4(self, first_word: int, v, elem=None):
        """OpMatrixTimesVector on a column-major mat4 at push-constant word `first_word` (SpirvShaderArithmetic.cpp:39-55):
        row i = M[i,0] * v0, then MulAdd(M[i,j], vj, .) for j = 1..3.  `elem(row, column)` names the operand that holds an element
        when the matrix lives somewhere else (a uniform block, a row-major layout)."""
        elem = elem or (lambda i, j: _push(first_word + 4 * j + i))
        out = []
        for i in range(4):
            acc = self.mul(elem(i, 0), v[0])
            for j in range(1, 4):
                acc = self.fma(elem(i, j), v[j], acc)
            out.append(acc)
        return out


def _vs(position, outputs: dict, program: _Program | None = None, point_size=None) -> capi.ShaderInfo:
    s = capi.ShaderInfo()
    s.stage = 0
    inmask = 0
    for i, (k, v) in enumerate(position):
        s.position[i] = capi.ShaderOperand(k, v)
        if k == capi.SRC_INPUT:
            inmask |= 1 << v
    for comp, (k, v) in outputs.items():
        s.output[comp] = capi.ShaderOperand(k, v)
        s.outputMask |= 1 << comp
        if k == capi.SRC_INPUT:
            inmask |= 1 << v
    if program is not None:
        s.programLength = len(program.steps)
        for i, (op, a, b, c) in enumerate(program.steps):
            s.program[i] = capi.ShaderOp(op, capi.ShaderOperand(*a), capi.ShaderOperand(*b), capi.ShaderOperand(*c))
            for (k, v) in (a, b, c):
                if k == capi.SRC_INPUT:
                    inmask |= 1 << v
    if point_size is not None:
        s.writesPointSize = 1
        s.pointSize = capi.ShaderOperand(*point_size)
        if point_size[0] == capi.SRC_INPUT:
            inmask |= 1 << point_size[1]
    s.inputMask = inmask
    return s


def _vs_mvp():
    # gl_Position = pc.mvp * vec4(inPos, 1.0); outColor = inColor
    p = _Program()
    pos = p.mat4_times_vec4(0, [_inp(0, 0), _inp(0, 1), _inp(0, 2), _const(1.0)])
    return _vs(pos, {0: _inp(1, 0), 1: _inp(1, 1), 2: _inp(1, 2), 3: _inp(1, 3)}, p)


def _vs_ubo():
    # layout(set = 0, binding = 1) uniform UBO { mat4 model (ColMajor, offset 0); row_major mat4 viewProj (offset 64); vec4 tint (offset 128); } u;
    # gl_Position = u.viewProj * (u.model * vec4(inPos, 1.0)); outColor = inColor * u.tint
    p = _Program()
    world = p.mat4_times_vec4(0, [_inp(0, 0), _inp(0, 1), _inp(0, 2), _const(1.0)], elem=lambda i, j: _ubo(0, 4 * j + i))
    clip = p.mat4_times_vec4(0, world, elem=lambda i, j: _ubo(0, 16 + 4 * i + j))  # row-major: the rows lie MatrixStride apart
    col = [p.mul(_inp(1, c), _ubo(0, 32 + c)) for c in range(4)]
    s = _vs(clip, {c: col[c] for c in range(4)}, p)
    s.uniformCount, s.uniformSet[0], s.uniformBinding[0] = 1, 0, 1
    return s


def _vs_inst():
    p = _Program()
    q = [p.add(_inp(0, c), _inp(2, c)) for c in range(4)]
    return _vs(q, {c: _inp(1, c) for c in range(4)}, p)


def _vs_xform():
    # p = inPos * pc.scale.xyz + pc.offset.xyz; gl_Position = vec4(p, pc.offset.w); outColor = inColor * pc.tint
    p = _Program()
    m = [p.mul(_inp(0, c), _push(c)) for c in range(3)]
    q = [p.add(m[c], _push(4 + c)) for c in range(3)]
    col = [p.mul(_inp(1, c), _push(8 + c)) for c in range(4)]
    return _vs(q + [_push(7)], {c: col[c] for c in range(4)}, p)


def _fs(colour, tex=None) -> capi.ShaderInfo:
    s = capi.ShaderInfo()
    s.stage = 4
    inmask = 0
    for i, (k, v) in enumerate(colour):
        s.output[i] = capi.ShaderOperand(k, v)
        s.outputMask |= 1 << i
        if k == capi.SRC_INPUT:
            inmask |= 1 << v
    if tex is not None:
        (set_, binding, u, v) = tex
        s.usesTexture, s.textureSet, s.textureBinding = 1, set_, binding
        for i, (k, val) in enumerate((u, v)):
            s.texCoord[i] = capi.ShaderOperand(k, val)
            if k == capi.SRC_INPUT:
                inmask |= 1 << val
    s.inputMask = inmask
    return s


# Hand-written meaning of each fixture shader (NOT derived from the SPIR-V).
SHADER_SPECS = {
    # gl_Position = vec4(inPos.xyz, 1.0)
    "vs_pos3": lambda: _vs([_inp(0, 0), _inp(0, 1), _inp(0, 2), _const(1.0)], {}),
    # outColor(loc0).rgb = inColor(loc1).rgb
    "vs_pos3_col3": lambda: _vs([_inp(0, 0), _inp(0, 1), _inp(0, 2), _const(1.0)],
                                {0: _inp(1, 0), 1: _inp(1, 1), 2: _inp(1, 2)}),
    # outColor(loc0) = inColor(loc1) (vec4)
    "vs_pos3_col4": lambda: _vs([_inp(0, 0), _inp(0, 1), _inp(0, 2), _const(1.0)],
                                {0: _inp(1, 0), 1: _inp(1, 1), 2: _inp(1, 2), 3: _inp(1, 3)}),
    # outTexCoord(loc0).xy = inTexCoord(loc1).xy
    "vs_pos3_uv2": lambda: _vs([_inp(0, 0), _inp(0, 1), _inp(0, 2), _const(1.0)], {0: _inp(1, 0), 1: _inp(1, 1)}),
    # gl_Position = inPos (vec4); outCol = inCol (vec4)
    "vs_pos4_col4": lambda: _vs([_inp(0, 0), _inp(0, 1), _inp(0, 2), _inp(0, 3)],
                                {0: _inp(1, 0), 1: _inp(1, 1), 2: _inp(1, 2), 3: _inp(1, 3)}),
    # gl_Position = pc.mvp * vec4(inPos, 1.0) with a mat4 in the push-constant block
    "vs_mvp_pos3_col4": _vs_mvp,
    # component-wise scale / offset / tint from the push-constant block
    "vs_xform_pos3_col4": _vs_xform,
    # two matrices (one row-major) and a tint from a uniform buffer at (set 0, binding 1)
    "vs_ubo_pos3_col4": _vs_ubo,
    # gl_Position = inPos; gl_PointSize = inSize (loc 2); outColor = inColor
    "vs_point_pos4_col4": lambda: _vs([_inp(0, 0), _inp(0, 1), _inp(0, 2), _inp(0, 3)],
                                      {0: _inp(1, 0), 1: _inp(1, 1), 2: _inp(1, 2), 3: _inp(1, 3)}, point_size=_inp(2, 0)),
    # gl_Position = inPos + instOffset (loc 2, per instance); outColor = instColor (loc 1, per instance)
    "vs_inst_pos4_col4": _vs_inst,
    "fs_white": lambda: _fs([_const(1.0)] * 4),
    "fs_col3": lambda: _fs([_inp(0, 0), _inp(0, 1), _inp(0, 2), _const(1.0)]),
    "fs_col4": lambda: _fs([_inp(0, 0), _inp(0, 1), _inp(0, 2), _inp(0, 3)]),
    "fs_tex_uv2": lambda: _fs([_texel(0), _texel(1), _texel(2), _texel(3)], tex=(0, 0, _inp(0, 0), _inp(0, 1))),
    "fs_tex_col4": lambda: _fs([_texel(0), _texel(1), _texel(2), _texel(3)], tex=(0, 0, _inp(0, 0), _inp(0, 1))),
}


def shader_spec(name: str) -> capi.ShaderInfo:
    return SHADER_SPECS[name]()


def render_oracle(scene: Scene, att: dict | None = None, render_area=None) -> dict:
    """Render every draw of the scene with the C restatement, in order, into host numpy attachments."""
    L = lib()
    att = att if att is not None else scene.alloc_attachments()
    for dr in scene.flat_draws():
        keep: list = []
        d = scene.build_desc(dr, att, keep, render_area)
        vs, fs = shader_spec(dr.vs), shader_spec(dr.fs)
        rc = L.swref_draw(C.byref(d), C.byref(vs), C.byref(fs))
        if rc != 0:
            raise RuntimeError(f"swref_draw failed: {rc}")
    return att


def resolve_oracle(scene: Scene, att: dict) -> np.ndarray:
    H2, W = scene.padded_height(), scene.width
    out = np.zeros((H2, W, 4), dtype=scene.color_dtype())
    bpp = scene.color_bpp()
    src = capi.Attachment(att["color"].ctypes.data, scene.colorFormat, W * bpp, H2 * W * bpp, W, scene.height, 0)
    dst = capi.Attachment(out.ctypes.data, scene.colorFormat, W * bpp, H2 * W * bpp, W, scene.height, 0)
    rc = lib().swref_resolve(C.byref(src), scene.samples, C.byref(dst))
    if rc != 0:
        raise RuntimeError(f"swref_resolve failed: {rc}")
    return out


def reference_available() -> bool:
    return os.path.exists(REF_ICD) and os.path.exists(REFRENDER)


def cuda_icd_available() -> bool:
    return os.path.exists(CUDA_ICD) and os.path.exists(REFRENDER)


def render_reference(scene: Scene, time_frames: int = 0, threads: int | None = None, warmup: int = 1, icd: str | None = None, env: dict | None = None) -> dict:
    """Render the scene through the Vulkan API of an ICD — by default the REFERENCE ICD (oracle/_ref); `icd` = CUDA_ICD drives the
    patched ICD instead.  Returns colour (resolved when multisampled), depth/stencil when single-sampled, and the timing JSON when
    time_frames > 0."""
    if icd is None and not reference_available():
        raise RuntimeError("reference ICD not built (oracle/build_ref.sh); only available where /root/reference is")
    with tempfile.TemporaryDirectory() as td:
        sp, op = os.path.join(td, "scene.bin"), os.path.join(td, "out.bin")
        scene.write_ref_scene(sp)
        if threads is not None:  # docs/RuntimeConfiguration.md:9-22 — SwiftShader.ini is read from the CWD
            with open(os.path.join(td, "SwiftShader.ini"), "w") as f:
                f.write(f"[Processor]\nThreadCount={threads}\n")
        cmd = [REFRENDER, icd or REF_ICD, sp, op]
        if time_frames:
            cmd += ["--time", str(time_frames), "--warmup", str(warmup)]
        res = subprocess.run(cmd, cwd=td, capture_output=True, text=True, env=None if env is None else {**os.environ, **env})
        if res.returncode != 0:
            raise RuntimeError(f"refrender failed ({res.returncode}): {res.stderr[-2000:]}")
        raw = open(op, "rb").read()
    magic, W, H, hasD, hasS, samples = struct.unpack_from("<6I", raw, 0)
    assert magic == 0x4F525753 and W == scene.width and H == scene.height
    off = 24
    cdt = scene.color_dtype()
    out = {"color": np.frombuffer(raw, cdt, W * H * 4, off).reshape(H, W, 4).copy()}
    off += W * H * 4 * cdt.itemsize
    if hasD == 2:  # D16_UNORM
        out["depth"] = np.frombuffer(raw, np.uint16, W * H, off).reshape(H, W).copy()
        off += W * H * 2
    elif hasD:
        out["depth"] = np.frombuffer(raw, np.float32, W * H, off).reshape(H, W).copy()
        off += W * H * 4
    if hasS:
        out["stencil"] = np.frombuffer(raw, np.uint8, W * H, off).reshape(H, W).copy()
    if time_frames:
        out["timing"] = json.loads(res.stdout.strip().splitlines()[-1])
    return out
