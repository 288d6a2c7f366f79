#!/bin/bash
# Builds the reference ICD WITH the CUDA draw path wired in behind sw::Renderer::draw:
#   oracle/_cuda/libvk_swiftshader_cuda.so  =  the object files of the unmodified ICD build (oracle/build_ref.sh, $SS_BUILD_DIR)
#                                            with Renderer.cpp, VkDeviceMemory.cpp and VkImageView.cpp replaced by patched copies (icd/swiftshader_cuda.patch)
#                                            plus icd/swcu_shim.cpp.
# Nothing is written under /root/reference and no reference source is copied into this repository: the three files are copied to a
# scratch directory, patched there, compiled with the very command lines of the reference's own build (ninja -t commands) and linked
# with the reference's own link line.  libswcuda.so is NOT linked: the shim dlopens it (SWCU_LIB, or ../../swiftshader_b200/csrc/).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REPO="$(cd "$HERE/.." && pwd)"
REF="${REF:-/root/reference}"
BUILD="${SS_BUILD_DIR:-/tmp/ss-build}"
SCRATCH="${SS_CUDA_SCRATCH:-/tmp/ss-cuda}"
OUT="$HERE/_cuda"
if [ ! -d "$REF" ]; then echo "no reference tree at $REF; cannot build the patched ICD" >&2; exit 0; fi
if [ ! -f "$BUILD/build.ninja" ]; then echo "no ICD build tree at $BUILD: run oracle/build_ref.sh first" >&2; exit 1; fi
mkdir -p "$OUT" "$SCRATCH/src/Device" "$SCRATCH/src/Vulkan" "$SCRATCH/obj"
cp "$REF/src/Device/Renderer.cpp" "$SCRATCH/src/Device/Renderer.cpp"
cp "$REF/src/Vulkan/VkDeviceMemory.cpp" "$SCRATCH/src/Vulkan/VkDeviceMemory.cpp"
cp "$REF/src/Vulkan/VkImageView.cpp" "$SCRATCH/src/Vulkan/VkImageView.cpp"
(cd "$SCRATCH" && patch -s -p1 < "$REPO/icd/swiftshader_cuda.patch")
cd "$BUILD"
ninja -t commands vk_swiftshader > "$SCRATCH/commands.txt"
# compile one (patched or new) source with the command line the reference build uses for `like`
compile() { # <like: source path in the reference> <source to compile> <object out> <extra -I dir>
  local cmd
  cmd="$(grep -F -- "-c $1" "$SCRATCH/commands.txt" | head -1)"
  [ -n "$cmd" ] || { echo "no compile command for $1" >&2; exit 1; }
  cmd="${cmd// -c $1/ -c $2}"
  cmd="$(echo "$cmd" | sed -E "s# -o [^ ]+# -o $3#; s# -MF [^ ]+# -MF $3.d#; s# -MT [^ ]+# -MT $3#")"
  cmd="${cmd/ -c / -I$4 -I$REPO/include -I$REPO/icd -I$REF/src -c }"
  eval "$cmd"
}
compile "$REF/src/Device/Renderer.cpp" "$SCRATCH/src/Device/Renderer.cpp" "$SCRATCH/obj/Renderer.cpp.o" "$REF/src/Device"
compile "$REF/src/Vulkan/VkDeviceMemory.cpp" "$SCRATCH/src/Vulkan/VkDeviceMemory.cpp" "$SCRATCH/obj/VkDeviceMemory.cpp.o" "$REF/src/Vulkan"
compile "$REF/src/Vulkan/VkImageView.cpp" "$SCRATCH/src/Vulkan/VkImageView.cpp" "$SCRATCH/obj/VkImageView.cpp.o" "$REF/src/Vulkan"
compile "$REF/src/Vulkan/VkDeviceMemory.cpp" "$REPO/icd/swcu_shim.cpp" "$SCRATCH/obj/swcu_shim.cpp.o" "$REF/src/Vulkan"
# libvk_device.a with the patched Renderer
cp "$BUILD/src/Device/libvk_device.a" "$SCRATCH/obj/libvk_device_cuda.a"
ar d "$SCRATCH/obj/libvk_device_cuda.a" Renderer.cpp.o
ar q "$SCRATCH/obj/libvk_device_cuda.a" "$SCRATCH/obj/Renderer.cpp.o"
# the reference's link line with the two objects swapped, the shim added, -ldl for dlopen
link="$(grep -F -- "-o libvk_swiftshader.so" "$SCRATCH/commands.txt" | head -1)"
link="${link#: && }"
link="${link%% && cd *}"   # (what follows only copies the library around the build tree)
link="${link//-o libvk_swiftshader.so/-o $OUT/libvk_swiftshader_cuda.so}"
link="${link//src\/Vulkan\/CMakeFiles\/vk_swiftshader.dir\/VkDeviceMemory.cpp.o/$SCRATCH/obj/VkDeviceMemory.cpp.o $SCRATCH/obj/swcu_shim.cpp.o}"
link="${link//src\/Vulkan\/CMakeFiles\/vk_swiftshader.dir\/VkImageView.cpp.o/$SCRATCH/obj/VkImageView.cpp.o}"
link="${link//src\/Device\/libvk_device.a/$SCRATCH/obj/libvk_device_cuda.a}"
link="$(echo "$link" | sed -E 's# -Wl,--dependency-file=[^ ]+##')"
eval "$link -ldl"
echo "built $OUT/libvk_swiftshader_cuda.so"
