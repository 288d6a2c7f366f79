"""Minimal SPIR-V text assembler for the benchmark shader subset.

There is no glslang / spirv-as in the image, so the shaders of the reference's triangle benchmarks
(/root/reference/tests/VulkanBenchmarks/TriangleBenchmarks.cpp:56-81,109-140,168-198) are kept as SPIR-V
assembly text under ``swiftshader_b200/shaders/*.spvasm`` and assembled here into the binary words that
``vkCreateShaderModule`` / ``swcu_draw_desc.vertexShader`` take.  The syntax is the standard
``%id = OpName operands`` form (what ``spirv-as`` accepts, cf. tests/VulkanUnitTests/ComputeTests.cpp:70-77);
only the opcodes below are known.  The *consumer* of these binaries on the product path is the C++ translator
in ``csrc/spirv_subset.cpp`` — this module only produces inputs.
"""
from __future__ import annotations

import os
import re
import struct

import numpy as np

MAGIC = 0x07230203
VERSION_1_3 = 0x00010300

ENUMS = {
    "Capability": {"Matrix": 0, "Shader": 1, "Sampled1D": 43, "ImageQuery": 50, "DerivativeControl": 51},
    "AddressingModel": {"Logical": 0},
    "MemoryModel": {"Simple": 0, "GLSL450": 1},
    "ExecutionModel": {"Vertex": 0, "Fragment": 4, "GLCompute": 5},
    "ExecutionMode": {"OriginUpperLeft": 7, "EarlyFragmentTests": 9, "DepthReplacing": 12},
    "StorageClass": {"UniformConstant": 0, "Input": 1, "Uniform": 2, "Output": 3, "Private": 6, "Function": 7,
                     "PushConstant": 9},
    "Dim": {"1D": 0, "2D": 1, "3D": 2, "Cube": 3},
    "ImageFormat": {"Unknown": 0, "Rgba8": 4},
    "FunctionControl": {"None": 0},
    "Decoration": {"RelaxedPrecision": 0, "Block": 2, "RowMajor": 4, "ColMajor": 5, "MatrixStride": 7, "BuiltIn": 11, "NoPerspective": 13, "Flat": 14,
                   "Centroid": 16, "Location": 30, "Component": 31, "Binding": 33, "DescriptorSet": 34, "Offset": 35},
    "BuiltIn": {"Position": 0, "PointSize": 1, "ClipDistance": 3, "CullDistance": 4, "FragCoord": 15,
                "FrontFacing": 17, "FragDepth": 22},
    "SourceLanguage": {"Unknown": 0, "ESSL": 1, "GLSL": 2},
    "ImageOperands": {"None": 0, "Bias": 1, "Lod": 2},
}

# name -> (opcode, has_result_type, has_result, operand kinds)
# kinds: id, lit, str, <EnumName>, ids* (rest are ids), lits* (rest literals), const (typed literal), deco (decoration extras)
OPS = {
    "OpNop": (0, False, False, []),
    "OpSource": (3, False, False, ["SourceLanguage", "lit"]),
    "OpName": (5, False, False, ["id", "str"]),
    "OpMemberName": (6, False, False, ["id", "lit", "str"]),
    "OpExtInstImport": (11, False, True, ["str"]),
    "OpExtInst": (12, True, True, ["id", "lit", "ids*"]),
    "OpMemoryModel": (14, False, False, ["AddressingModel", "MemoryModel"]),
    "OpEntryPoint": (15, False, False, ["ExecutionModel", "id", "str", "ids*"]),
    "OpExecutionMode": (16, False, False, ["id", "ExecutionMode", "lits*"]),
    "OpCapability": (17, False, False, ["Capability"]),
    "OpTypeVoid": (19, False, True, []),
    "OpTypeBool": (20, False, True, []),
    "OpTypeInt": (21, False, True, ["lit", "lit"]),
    "OpTypeFloat": (22, False, True, ["lit"]),
    "OpTypeVector": (23, False, True, ["id", "lit"]),
    "OpTypeMatrix": (24, False, True, ["id", "lit"]),
    "OpTypeImage": (25, False, True, ["id", "Dim", "lit", "lit", "lit", "lit", "ImageFormat"]),
    "OpTypeSampler": (26, False, True, []),
    "OpTypeSampledImage": (27, False, True, ["id"]),
    "OpTypeArray": (28, False, True, ["id", "id"]),
    "OpTypeStruct": (30, False, True, ["ids*"]),
    "OpTypePointer": (32, False, True, ["StorageClass", "id"]),
    "OpTypeFunction": (33, False, True, ["id", "ids*"]),
    "OpConstant": (43, True, True, ["const"]),
    "OpConstantComposite": (44, True, True, ["ids*"]),
    "OpFunction": (54, True, True, ["FunctionControl", "id"]),
    "OpFunctionEnd": (56, False, False, []),
    "OpVariable": (59, True, True, ["StorageClass", "ids*"]),
    "OpLoad": (61, True, True, ["id"]),
    "OpStore": (62, False, False, ["id", "id"]),
    "OpAccessChain": (65, True, True, ["id", "ids*"]),
    "OpDecorate": (71, False, False, ["id", "deco"]),
    "OpMemberDecorate": (72, False, False, ["id", "lit", "deco"]),
    "OpVectorShuffle": (79, True, True, ["id", "id", "lits*"]),
    "OpCompositeConstruct": (80, True, True, ["ids*"]),
    "OpCompositeExtract": (81, True, True, ["id", "lits*"]),
    "OpCompositeInsert": (82, True, True, ["id", "id", "lits*"]),
    "OpCopyObject": (83, True, True, ["id"]),
    "OpImageSampleImplicitLod": (87, True, True, ["id", "id", "lits*"]),
    "OpImageSampleExplicitLod": (88, True, True, ["id", "id", "ImageOperands", "ids*"]),
    "OpFNegate": (127, True, True, ["id"]),
    "OpFAdd": (129, True, True, ["id", "id"]),
    "OpFSub": (131, True, True, ["id", "id"]),
    "OpFMul": (133, True, True, ["id", "id"]),
    "OpVectorTimesScalar": (142, True, True, ["id", "id"]),
    "OpMatrixTimesScalar": (143, True, True, ["id", "id"]),
    "OpVectorTimesMatrix": (144, True, True, ["id", "id"]),
    "OpMatrixTimesVector": (145, True, True, ["id", "id"]),
    "OpDot": (148, True, True, ["id", "id"]),
    "OpLabel": (248, False, True, []),
    "OpKill": (252, False, False, []),
    "OpReturn": (253, False, False, []),
}

_TOKEN = re.compile(r'"(?:[^"\\]|\\.)*"|[^\s]+')


def _encode_string(s: str) -> list[int]:
    b = s.encode("utf-8") + b"\0"
    b += b"\0" * ((-len(b)) % 4)
    return list(struct.unpack("<%dI" % (len(b) // 4), b))


def assemble(text: str) -> np.ndarray:
    """Assemble SPIR-V assembly text into a uint32 word array (little-endian module)."""
    ids: dict[str, int] = {}
    float_types: dict[int, int] = {}  # type id -> width

    def get_id(tok: str) -> int:
        if not tok.startswith("%"):
            raise ValueError(f"expected %id, got {tok!r}")
        if tok not in ids:
            ids[tok] = len(ids) + 1
        return ids[tok]

    def lit(tok: str) -> int:
        return int(tok, 0) & 0xFFFFFFFF

    words: list[int] = []
    for lineno, raw in enumerate(text.splitlines(), 1):
        line = raw.split(";", 1)[0].strip()
        if not line:
            continue
        toks = _TOKEN.findall(line)
        result = None
        if len(toks) >= 3 and toks[1] == "=":
            result = toks[0]
            toks = toks[2:]
        name = toks[0]
        if name not in OPS:
            raise ValueError(f"line {lineno}: unknown opcode {name}")
        opcode, has_type, has_result, kinds = OPS[name]
        args = toks[1:]
        out: list[int] = []
        type_id = None
        if has_type:
            type_id = get_id(args.pop(0))
            out.append(type_id)
        if has_result:
            if result is None:
                raise ValueError(f"line {lineno}: {name} needs a result id")
            out.append(get_id(result))
        for kind in kinds:
            if kind == "ids*":
                out += [get_id(a) for a in args]
                args = []
            elif kind == "lits*":
                out += [lit(a) for a in args]
                args = []
            elif kind == "deco":
                d = args.pop(0)
                out.append(ENUMS["Decoration"][d])
                if d == "BuiltIn":
                    out.append(ENUMS["BuiltIn"][args.pop(0)])
                out += [lit(a) for a in args]
                args = []
            elif not args:
                raise ValueError(f"line {lineno}: missing operand {kind} for {name}")
            elif kind == "id":
                out.append(get_id(args.pop(0)))
            elif kind == "lit":
                out.append(lit(args.pop(0)))
            elif kind == "str":
                out += _encode_string(args.pop(0)[1:-1])
            elif kind == "const":
                a = args.pop(0)
                if type_id in float_types:
                    out.append(struct.unpack("<I", struct.pack("<f", float(a)))[0])
                else:
                    out.append(lit(a))
            elif kind in ENUMS:
                a = args.pop(0)
                table = ENUMS[kind]
                if a in table:
                    out.append(table[a])
                elif "|" in a:
                    v = 0
                    for part in a.split("|"):
                        v |= table[part]
                    out.append(v)
                else:
                    out.append(lit(a))
            else:
                raise AssertionError(kind)
        if args:
            raise ValueError(f"line {lineno}: extra operands {args} for {name}")
        if name == "OpTypeFloat":
            float_types[out[0]] = out[1]
        words.append(((len(out) + 1) << 16) | opcode)
        words += out
    header = [MAGIC, VERSION_1_3, 0, len(ids) + 1, 0]
    return np.array(header + words, dtype=np.uint32)


_SHADER_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shaders")
_cache: dict[str, np.ndarray] = {}


def shader_source(name: str) -> str:
    with open(os.path.join(_SHADER_DIR, name + ".spvasm")) as f:
        return f.read()


def shader(name: str) -> np.ndarray:
    """Assembled words of the fixture shader ``swiftshader_b200/shaders/<name>.spvasm``."""
    if name not in _cache:
        _cache[name] = assemble(shader_source(name))
    return _cache[name]
