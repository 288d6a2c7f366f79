"""CPU: the C-ABI library loads and exports every symbol include/swcu.h declares; the ctypes mirror has the same struct
sizes as the C compiler; the SPIR-V subset translator accepts the fixture shaders and rejects everything else; error
behaviour without a GPU is loud (no fallback)."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from oracle import swref
from swiftshader_b200 import capi, spirv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_match_header():
    hdr = open(os.path.join(ROOT, "include", "swcu.h")).read()
    declared = set(re.findall(r"\b(swcu_[a-z_]+)\s*\(", hdr))
    assert declared == set(capi.EXPORTS)
    lib = capi.lib()
    for name in declared:
        assert getattr(lib, name) is not None


def test_struct_sizes_match_c():
    src = r'''
#include "swcu.h"
#include <stdio.h>
int main(){ printf("%zu %zu %zu %zu %zu %zu\n", sizeof(swcu_draw_desc), sizeof(swcu_sampled_image), sizeof(swcu_attachment),
  sizeof(swcu_shader_info), sizeof(swcu_stats), sizeof(swcu_vertex_input)); return 0; }'''
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "s.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(td, "s"), os.path.join(td, "s.c")])
        got = [int(x) for x in subprocess.check_output([os.path.join(td, "s")]).split()]
    want = [C.sizeof(t) for t in (capi.DrawDesc, capi.SampledImage, capi.Attachment, capi.ShaderInfo, capi.Stats, capi.VertexInput)]
    assert got == want


def _dump(i):
    return dict(stage=i.stage, outputMask=i.outputMask, position=[(o.kind, o.value) for o in i.position],
                output=[(o.kind, o.value) for o in i.output], inputMask=i.inputMask, flat=i.flatMask, nopersp=i.noPerspectiveMask,
                tex=(i.usesTexture, i.textureSet, i.textureBinding), tc=[(o.kind, o.value) for o in i.texCoord],
                program=[(s.op, (s.a.kind, s.a.value), (s.b.kind, s.b.value), (s.c.kind, s.c.value)) for s in i.program[:i.programLength]],
                pointSize=(i.writesPointSize, i.pointSize.kind, i.pointSize.value),
                uniforms=[(i.uniformSet[k], i.uniformBinding[k]) for k in range(i.uniformCount)])


@pytest.mark.parametrize("name", sorted(swref.SHADER_SPECS))
def test_translator_matches_hand_written_meaning(name):
    assert _dump(capi.translate_shader(spirv.shader(name))) == _dump(swref.shader_spec(name))


@pytest.mark.parametrize("name", sorted(swref.SHADER_SPECS))
def test_translator_on_the_post_spirv_opt_form(name):
    """The shim hands over SpirvShader::insns, i.e. the module AFTER spirv-opt (VkPipeline.cpp:36-107); the committed fixtures are the
    fixture shaders run through the reference's own pass list (tests/golden/gen_spv_postopt.py).  Same meaning either way."""
    words = np.fromfile(os.path.join(ROOT, "tests", "golden", "spv_postopt", name + ".spv"), dtype=np.uint32)
    assert _dump(capi.translate_shader(words)) == _dump(swref.shader_spec(name))


def test_translator_vertex_arithmetic_limits():
    """a push-constant access beyond 128 bytes, and arithmetic in the fragment stage, are outside the subset"""
    src = spirv.shader_source("vs_mvp_pos3_col4").replace("OpMemberDecorate %PC 0 Offset 0", "OpMemberDecorate %PC 0 Offset 96")
    with pytest.raises(capi.SwcuError) as e:
        capi.translate_shader(spirv.assemble(src))
    assert e.value.code == capi.E_UNSUPPORTED and "push-constant access beyond" in str(e.value)


def test_translator_uniform_buffer_limits():
    """a uniform block needs DescriptorSet / Binding and the Block decoration; words past 64 KiB are outside the subset"""
    src = spirv.shader_source("vs_ubo_pos3_col4")
    info = capi.translate_shader(spirv.assemble(src))
    assert info.uniformCount == 1 and (info.uniformSet[0], info.uniformBinding[0]) == (0, 1)
    assert info.program[16].a.kind == capi.SRC_UNIFORM and info.program[16].a.value == 16  # row-major viewProj: element (0, 0) at byte 64
    assert info.program[17].a.value == 17 and info.program[20].a.value == 20              # (0, 1) 4 bytes on, (1, 0) one MatrixStride on
    for bad, what in ((src.replace("OpDecorate %u Binding 1\n", ""), "without DescriptorSet/Binding"),
                      (src.replace("OpDecorate %UBO Block\n", ""), "must be a Block struct"),
                      (src.replace("OpMemberDecorate %UBO 2 Offset 128", "OpMemberDecorate %UBO 2 Offset 65536"), "uniform-buffer access beyond")):
        with pytest.raises(capi.SwcuError) as e:
            capi.translate_shader(spirv.assemble(bad))
        assert e.value.code == capi.E_UNSUPPORTED and what in str(e.value)


FS_HEAD = """OpCapability Shader
OpMemoryModel Logical GLSL450
OpEntryPoint Fragment %main "main" %outColor %inColor
OpExecutionMode %main OriginUpperLeft
OpDecorate %outColor Location 0
OpDecorate %inColor Location 0
{deco}
%void = OpTypeVoid
%fn = OpTypeFunction %void
%float = OpTypeFloat 32
%v4 = OpTypeVector %float 4
%ptr_out_v4 = OpTypePointer Output %v4
%outColor = OpVariable %ptr_out_v4 Output
%ptr_in_v4 = OpTypePointer Input %v4
%inColor = OpVariable %ptr_in_v4 Input
%main = OpFunction %void None %fn
%l = OpLabel
%c = OpLoad %v4 %inColor
{body}
OpReturn
OpFunctionEnd
"""


def test_translator_qualifiers_and_shuffle():
    info = capi.translate_shader(spirv.assemble(FS_HEAD.format(deco="OpDecorate %inColor Flat",
                                                               body="%s = OpVectorShuffle %v4 %c %c 2 1 0 7\nOpStore %outColor %s")))
    assert [(o.kind, o.value) for o in info.output][:4] == [(0, 2), (0, 1), (0, 0), (0, 3)]
    assert info.flatMask == 0xF and info.noPerspectiveMask == 0 and info.inputMask == 0xF
    info = capi.translate_shader(spirv.assemble(FS_HEAD.format(deco="OpDecorate %inColor NoPerspective", body="OpStore %outColor %c")))
    assert info.noPerspectiveMask == 0xF


@pytest.mark.parametrize("body,what", [
    ("%s = OpFAdd %v4 %c %c\nOpStore %outColor %s", "opcode 129"),
    ("%s = OpFMul %v4 %c %c\nOpStore %outColor %s", "opcode 133"),
    ("OpKill", "opcode 252"),
    ("", "does not write colour"),
])
def test_translator_rejects_outside_subset(body, what):
    with pytest.raises(capi.SwcuError) as e:
        capi.translate_shader(spirv.assemble(FS_HEAD.format(deco="", body=body)))
    assert e.value.code == capi.E_UNSUPPORTED
    assert what in str(e.value)


def test_translator_rejects_garbage():
    with pytest.raises(capi.SwcuError) as e:
        capi.translate_shader(np.zeros(8, dtype=np.uint32))
    assert e.value.code == capi.E_INVALID
    compute = spirv.assemble("""OpCapability Shader
OpMemoryModel Logical GLSL450
OpEntryPoint GLCompute %main "main"
%void = OpTypeVoid
%fn = OpTypeFunction %void
%main = OpFunction %void None %fn
%l = OpLabel
OpReturn
OpFunctionEnd
""")
    with pytest.raises(capi.SwcuError) as e:
        capi.translate_shader(compute)
    assert e.value.code == capi.E_UNSUPPORTED


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    ctx = C.c_void_p()
    rc = capi.lib().swcu_create(C.byref(ctx), 0)
    assert rc == capi.E_CUDA and not ctx
    assert b"no CPU fallback" in capi.lib().swcu_last_error(None)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "swiftshader_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, re.M), f
                assert "swref" not in text and "_ref/" not in text.replace("oracle/_ref", ""), f
