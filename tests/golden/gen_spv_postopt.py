"""Writes tests/golden/spv_postopt/<shader>.spv: every fixture shader of swiftshader_b200/shaders/ in the form the reference's pipeline
holds it — after spirv-opt (src/Vulkan/VkPipeline.cpp:36-107: CreateRemoveDontInlinePass + RegisterPerformancePasses), produced by
oracle/_ref/spvopt (oracle/spvopt.cpp linked against the SPIRV-Tools of the reference build).  Runs only where /root/reference and its
build tree are; the .spv files are committed so that tests/test_boundary.py can translate both forms anywhere.
usage: python tests/golden/gen_spv_postopt.py"""
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..", "..")
sys.path.insert(0, ROOT)
from swiftshader_b200 import spirv  # noqa: E402

TOOL = os.path.join(ROOT, "oracle", "_ref", "spvopt")
OUT = os.path.join(HERE, "spv_postopt")


def main():
    os.makedirs(OUT, exist_ok=True)
    names = sorted(f[:-7] for f in os.listdir(os.path.join(ROOT, "swiftshader_b200", "shaders")) if f.endswith(".spvasm"))
    with tempfile.TemporaryDirectory() as td:
        for n in names:
            src = os.path.join(td, n + ".spv")
            spirv.shader(n).tofile(src)
            subprocess.check_call([TOOL, src, os.path.join(OUT, n + ".spv")])
            print(n, os.path.getsize(src), "->", os.path.getsize(os.path.join(OUT, n + ".spv")), "bytes")


if __name__ == "__main__":
    main()
