"""GPU: the CUDA draw path THROUGH THE VULKAN API — the reference ICD rebuilt with icd/swiftshader_cuda.patch
(oracle/build_cuda_icd.sh -> oracle/_cuda/libvk_swiftshader_cuda.so): vkCmdDraw recording, CmdDrawBase::draw and the state gathering
of sw::Renderer::draw are the reference's own code, DrawCall::run is swcu_draw, vk::DeviceMemory allocations have device shadows.
The same harness that renders the goldens with the unmodified ICD (oracle/refrender.cpp) drives it; the outputs must have the
goldens' hashes.  This is also where the translator sees the shaders in the form the pipeline holds them — after spirv-opt
(src/Vulkan/VkPipeline.cpp:42-107) — instead of the hand-written fixtures."""
import hashlib
import json
import os

import numpy as np
import pytest

import scenes
from oracle import swref

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
HASHES = json.load(open(os.path.join(HERE, "golden", "golden_hashes.json")))
CASES = dict(scenes.all_cases())
# the three VulkanBenchmarks triangles (1x and 4x), the reduced BASELINE workloads, and a few scenes of every state family
NAMES = ["benchmark_0_720p", "benchmark_1_720p", "benchmark_2_720p", "benchmark_0_720p_msaa", "benchmark_2_720p_msaa",
         "workload_c1", "workload_c2", "workload_c3", "workload_c4", "workload_c5",
         "coverage_0", "coverage_7", "zclip_1", "cull_3", "depth_ops_5", "blend_4", "blend_17", "stencil_3", "stencil_10", "texture_2", "texture_9",
         "msaa_1", "msaa_4", "topology_1", "topology_4", "scissor_2", "depth16_3", "srgb_2", "floatrt_5", "fragtests_6", "texsplit_3", "blendoff_0",
         # f4: a transform from push constants in the vertex stage (vkCmdPushConstants -> DrawData::pushConstants), lines, points
         "mvp_0", "mvp_2", "mvp_4", "lines_1", "lines_2", "lines_3", "points_0", "points_2", "points_5",
         # depthClampEnable; 4x MSAA on the formats whose resolve is the generic blit
         "zclamp_1", "zclamp_3", "msaafmt_0", "msaafmt_1", "msaafmt_2", "msaafmt_3", "msaafmt_5",
         # instanced draws: the reference's own loop over the instances (CmdDrawBase::draw) calls the CUDA path once per instance
         "instanced_0", "instanced_1", "instanced_3",
         # a transform from a uniform buffer: the shim hands over the vk::BufferDescriptor of the vertex shader's block (with a texture: both descriptors of the set)
         "ubo_0", "ubo_1", "ubo_2", "ubo_3"]


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("name", NAMES)
def test_patched_icd_renders_the_goldens_through_vulkan(device, name):
    if not swref.cuda_icd_available():
        pytest.skip("oracle/_cuda/libvk_swiftshader_cuda.so not built (oracle/build_cuda_icd.sh, where /root/reference is)")
    out = swref.render_reference(CASES[name], icd=swref.CUDA_ICD, env={"SWCU_ICD": "1"})
    for k, h in HASHES[name].items():
        assert _sha(out[k]) == h, f"{name}/{k}: the patched ICD's render differs from the reference ICD's"


def test_patched_icd_really_runs_the_cuda_path(device):
    """A draw outside the subset must abort the patched ICD (no silent fallback to the reference's routines) ..."""
    if not swref.cuda_icd_available():
        pytest.skip("patched ICD not built")
    import dataclasses
    sc = scenes.benchmark(1, 64, 64)
    sc = dataclasses.replace(sc, colorFormat=64)  # A2B10G10R10_UNORM_PACK32
    with pytest.raises(RuntimeError) as e:
        swref.render_reference(sc, icd=swref.CUDA_ICD, env={"SWCU_ICD": "1"})
    assert "swiftshader-cuda" in str(e.value)
