"""ctypes binding of the C-ABI in include/swcu.h (libswcuda.so).

This is plumbing only: the structures mirror swcu.h field for field and the loader fails loudly when the
CUDA library has not been built — there is no CPU fallback on the product path.
"""
from __future__ import annotations

import ctypes as C
import os

MAX_INPUTS = 16
MAX_SAMPLED_IMAGES = 4
MIPMAP_LEVELS = 15
MAX_VARYING_COMPONENTS = 16

OK, E_UNSUPPORTED, E_INVALID, E_CUDA, E_NOMEM = 0, -1, -2, -3, -4


class VertexInput(C.Structure):
    _fields_ = [("buffer", C.c_void_p), ("robustnessSize", C.c_uint32), ("vertexStride", C.c_uint32),
                ("format", C.c_uint32), ("reserved", C.c_uint32)]


class MipLevel(C.Structure):
    _fields_ = [("buffer", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32), ("pitchP", C.c_uint32),
                ("reserved", C.c_uint32)]


class SampledImage(C.Structure):
    _fields_ = [("set", C.c_uint32), ("binding", C.c_uint32), ("format", C.c_uint32), ("levelCount", C.c_uint32),
                ("level", MipLevel * MIPMAP_LEVELS),
                ("magFilter", C.c_uint32), ("minFilter", C.c_uint32), ("mipmapMode", C.c_uint32),
                ("addressModeU", C.c_uint32), ("addressModeV", C.c_uint32),
                ("mipLodBias", C.c_float), ("minLod", C.c_float), ("maxLod", C.c_float),
                ("anisotropyEnable", C.c_uint32), ("compareEnable", C.c_uint32),
                ("unnormalizedCoordinates", C.c_uint32), ("reserved", C.c_uint32)]


class StencilFace(C.Structure):
    _fields_ = [("failOp", C.c_uint32), ("passOp", C.c_uint32), ("depthFailOp", C.c_uint32), ("compareOp", C.c_uint32),
                ("compareMask", C.c_uint32), ("writeMask", C.c_uint32), ("reference", C.c_uint32)]


class Attachment(C.Structure):
    _fields_ = [("buffer", C.c_void_p), ("format", C.c_uint32), ("pitchB", C.c_int32), ("sliceB", C.c_int32),
                ("width", C.c_uint32), ("height", C.c_uint32), ("reserved", C.c_uint32)]


class Rect(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("width", C.c_uint32), ("height", C.c_uint32)]


MAX_UNIFORM_BUFFERS = 4


class UniformBuffer(C.Structure):
    _fields_ = [("set", C.c_uint32), ("binding", C.c_uint32), ("data", C.c_void_p), ("bytes", C.c_uint32), ("reserved0", C.c_uint32)]


class DrawDesc(C.Structure):
    _fields_ = [
        ("structSize", C.c_uint32),
        ("topology", C.c_uint32), ("provokingVertexMode", C.c_uint32), ("indexType", C.c_uint32),
        ("indexBuffer", C.c_void_p), ("primitiveCount", C.c_uint32), ("baseVertex", C.c_int32),
        ("input", VertexInput * MAX_INPUTS),
        ("vertexShader", C.c_void_p), ("vertexShaderWords", C.c_uint32), ("fragmentShaderWords", C.c_uint32),
        ("fragmentShader", C.c_void_p),
        ("viewportX", C.c_float), ("viewportY", C.c_float), ("viewportWidth", C.c_float), ("viewportHeight", C.c_float),
        ("viewportMinDepth", C.c_float), ("viewportMaxDepth", C.c_float),
        ("scissor", Rect), ("renderArea", Rect),
        ("cullMode", C.c_uint32), ("frontFace", C.c_uint32), ("depthClipEnable", C.c_uint32), ("depthClampEnable", C.c_uint32),
        ("depthBiasConstant", C.c_float), ("depthBiasSlope", C.c_float), ("depthBiasClamp", C.c_float),
        ("sampleCount", C.c_uint32), ("sampleMask", C.c_uint32), ("alphaToCoverageEnable", C.c_uint32),
        ("depthTestEnable", C.c_uint32), ("depthWriteEnable", C.c_uint32), ("depthCompareOp", C.c_uint32),
        ("stencilTestEnable", C.c_uint32), ("front", StencilFace), ("back", StencilFace),
        ("depthBoundsTestEnable", C.c_uint32), ("minDepthBounds", C.c_float), ("maxDepthBounds", C.c_float),
        ("blendEnable", C.c_uint32),
        ("srcColorBlendFactor", C.c_uint32), ("dstColorBlendFactor", C.c_uint32), ("colorBlendOp", C.c_uint32),
        ("srcAlphaBlendFactor", C.c_uint32), ("dstAlphaBlendFactor", C.c_uint32), ("alphaBlendOp", C.c_uint32),
        ("colorWriteMask", C.c_uint32), ("blendConstants", C.c_float * 4),
        ("color", Attachment), ("depth", Attachment), ("stencil", Attachment),
        ("pushConstants", C.c_void_p), ("pushConstantBytes", C.c_uint32), ("lineWidth", C.c_float),
        ("sampledImageCount", C.c_uint32), ("uniformBufferCount", C.c_uint32),
        ("sampledImage", SampledImage * MAX_SAMPLED_IMAGES),
        ("uniformBuffer", UniformBuffer * MAX_UNIFORM_BUFFERS),
    ]


class ShaderOperand(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("value", C.c_uint32)]


SRC_INPUT, SRC_CONST, SRC_TEXEL, SRC_PUSH, SRC_TEMP, SRC_UNIFORM = 0, 1, 2, 3, 4, 5
OP_MUL, OP_ADD, OP_SUB, OP_FMA, OP_NEG = 0, 1, 2, 3, 4
MAX_PROGRAM = 48


class ShaderOp(C.Structure):
    _fields_ = [("op", C.c_uint32), ("a", ShaderOperand), ("b", ShaderOperand), ("c", ShaderOperand)]



class ShaderInfo(C.Structure):
    _fields_ = [("stage", C.c_uint32), ("outputMask", C.c_uint32),
                ("position", ShaderOperand * 4), ("output", ShaderOperand * MAX_VARYING_COMPONENTS),
                ("inputMask", C.c_uint32), ("flatMask", C.c_uint32), ("noPerspectiveMask", C.c_uint32),
                ("usesTexture", C.c_uint32), ("textureSet", C.c_uint32), ("textureBinding", C.c_uint32),
                ("texCoord", ShaderOperand * 2),
                ("programLength", C.c_uint32), ("program", ShaderOp * MAX_PROGRAM),
                ("writesPointSize", C.c_uint32), ("pointSize", ShaderOperand),
                ("uniformCount", C.c_uint32), ("uniformSet", C.c_uint32 * MAX_UNIFORM_BUFFERS), ("uniformBinding", C.c_uint32 * MAX_UNIFORM_BUFFERS)]


class Stats(C.Structure):
    _fields_ = [("draws", C.c_uint64), ("kernelLaunches", C.c_uint64), ("primitives", C.c_uint64),
                ("h2dBytes", C.c_uint64), ("d2hBytes", C.c_uint64), ("pairs", C.c_uint64)]


class GroupDesc(C.Structure):
    _fields_ = [("structSize", C.c_uint32), ("rank", C.c_uint32), ("world", C.c_uint32), ("maxPrimitives", C.c_uint32),
                ("maxSlots", C.c_uint32), ("maxSamples", C.c_uint32), ("fbWidth", C.c_uint32), ("fbHeight", C.c_uint32)]


EXPORTS = [
    "swcu_create", "swcu_destroy", "swcu_last_error",
    "swcu_mem_register", "swcu_mem_register_device", "swcu_mem_unregister", "swcu_mem_upload", "swcu_mem_download", "swcu_mem_device_ptr",
    "swcu_draw", "swcu_sync", "swcu_clear", "swcu_resolve", "swcu_shader_translate",
    "swcu_set_stream", "swcu_timer_begin", "swcu_timer_end", "swcu_get_stats", "swcu_reset_stats",
    "swcu_set_profiling", "swcu_last_draw_kernels", "swcu_timeline", "swcu_side_begin", "swcu_side_end", "swcu_side_wait", "swcu_set_option", "swcu_version",
    "swcu_ipc_export", "swcu_ipc_open", "swcu_ipc_close", "swcu_copy_image", "swcu_signal", "swcu_wait_flags",
    "swcu_fence_signal", "swcu_fence_wait",
    "swcu_group_reserve", "swcu_group_attach", "swcu_group_detach", "swcu_mem_acquire", "swcu_mem_release", "swcu_mem_acquire_on", "swcu_mem_release_on",
]

_ROOT = os.path.dirname(os.path.abspath(__file__))
# SWCU_LIB: a differently configured build of the same library (kernel tuning experiments: scripts/gpu_variants.sh)
LIB_PATH = os.environ.get("SWCU_LIB") or os.path.join(_ROOT, "csrc", "libswcuda.so")
_lib = None


def lib() -> C.CDLL:
    """Load csrc/libswcuda.so (built by __graft_entry__.build()). Raises if it is missing: no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the draw path has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, u32, sz, i32 = C.c_void_p, C.c_uint32, C.c_size_t, C.c_int
    L.swcu_create.argtypes = [C.POINTER(vp), i32]
    L.swcu_destroy.argtypes = [vp]
    L.swcu_destroy.restype = None
    L.swcu_last_error.argtypes = [vp]
    L.swcu_last_error.restype = C.c_char_p
    L.swcu_mem_register.argtypes = [vp, vp, sz]
    L.swcu_mem_register_device.argtypes = [vp, vp, sz]
    L.swcu_mem_unregister.argtypes = [vp, vp]
    L.swcu_mem_upload.argtypes = [vp, vp, sz]
    L.swcu_mem_download.argtypes = [vp, vp, sz]
    L.swcu_mem_device_ptr.argtypes = [vp, vp]
    L.swcu_mem_device_ptr.restype = vp
    L.swcu_draw.argtypes = [vp, C.POINTER(DrawDesc)]
    L.swcu_sync.argtypes = [vp]
    L.swcu_clear.argtypes = [vp, C.POINTER(Attachment), u32, C.POINTER(Rect), vp]
    L.swcu_resolve.argtypes = [vp, C.POINTER(Attachment), u32, C.POINTER(Attachment)]
    L.swcu_shader_translate.argtypes = [vp, u32, C.POINTER(ShaderInfo), C.c_char_p, sz]
    L.swcu_set_stream.argtypes = [vp, vp]
    L.swcu_timer_begin.argtypes = [vp]
    L.swcu_timer_end.argtypes = [vp, C.POINTER(C.c_float)]
    L.swcu_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.swcu_reset_stats.argtypes = [vp]
    L.swcu_set_profiling.argtypes = [vp, i32]
    L.swcu_last_draw_kernels.argtypes = [vp, C.POINTER(C.c_char_p), C.POINTER(C.c_float), i32]
    L.swcu_set_option.argtypes = [vp, C.c_char_p, i32]
    L.swcu_timeline.argtypes = [vp, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.POINTER(C.c_float), i32]
    L.swcu_ipc_export.argtypes = [vp, vp, vp, C.POINTER(C.c_uint64)]
    L.swcu_ipc_open.argtypes = [vp, vp, C.POINTER(vp)]
    L.swcu_ipc_close.argtypes = [vp, vp]
    L.swcu_copy_image.argtypes = [vp, C.POINTER(Attachment), C.POINTER(Attachment)]
    L.swcu_signal.argtypes = [vp, vp, u32]
    L.swcu_wait_flags.argtypes = [vp, vp, u32, u32, u32]
    L.swcu_side_begin.argtypes = [vp]
    L.swcu_side_end.argtypes = [vp, u32]
    L.swcu_side_wait.argtypes = [vp, u32]
    L.swcu_fence_signal.argtypes = [vp, u32]
    L.swcu_fence_wait.argtypes = [vp, u32]
    L.swcu_group_reserve.argtypes = [vp, C.POINTER(GroupDesc), vp]
    L.swcu_group_attach.argtypes = [vp, vp]
    L.swcu_group_detach.argtypes = [vp]
    L.swcu_mem_acquire.argtypes = [vp, vp]
    L.swcu_mem_release.argtypes = [vp, vp]
    L.swcu_mem_acquire_on.argtypes = [vp, vp, vp]
    L.swcu_mem_release_on.argtypes = [vp, vp, vp]
    L.swcu_version.argtypes = []
    L.swcu_version.restype = C.c_char_p
    for name in EXPORTS:
        fn = getattr(L, name)
        if fn.restype is C.c_int and name not in ("swcu_destroy",):
            fn.restype = C.c_int
    _lib = L
    return L


class SwcuError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"swcu error {code}: {msg}")
        self.code = code


def translate_shader(words) -> ShaderInfo:
    """Run the narrow SPIR-V translator (host code inside libswcuda.so; needs no GPU)."""
    import numpy as np
    w = np.ascontiguousarray(words, dtype=np.uint32)
    info = ShaderInfo()
    err = C.create_string_buffer(512)
    rc = lib().swcu_shader_translate(w.ctypes.data, len(w), C.byref(info), err, len(err))
    if rc != OK:
        raise SwcuError(rc, err.value.decode())
    return info
