#!/bin/bash
# Builds the UNMODIFIED reference ICD (libvk_swiftshader.so) out-of-tree from /root/reference into
# oracle/_ref/ (git-ignored; travels to the GPU box).  The draw path is JIT-generated through Reactor +
# the vendored LLVM 10, so it cannot be compiled "from a few source files": this is the one place the
# reference's CMake build is used (ICD target only; ~11 min on 8 cores).
# Nothing is written under /root/reference and no reference source is copied into this repo.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REF:-/root/reference}"
BUILD="${SS_BUILD_DIR:-/tmp/ss-build}"
OUT="$HERE/_ref"
mkdir -p "$OUT"
if [ -f "$OUT/libvk_swiftshader.so" ]; then echo "oracle/_ref/libvk_swiftshader.so already present"; exit 0; fi
if [ ! -d "$REF" ]; then echo "no reference tree at $REF; cannot build the reference ICD" >&2; exit 0; fi
if [ ! -f "$BUILD/Linux/libvk_swiftshader.so" ]; then
  mkdir -p "$BUILD"
  cmake -G Ninja -S "$REF" -B "$BUILD" -DCMAKE_BUILD_TYPE=Release \
    -DSWIFTSHADER_BUILD_TESTS=OFF -DSWIFTSHADER_BUILD_BENCHMARKS=OFF \
    -DSWIFTSHADER_BUILD_WSI_XCB=OFF -DSWIFTSHADER_BUILD_WSI_WAYLAND=OFF \
    -DSWIFTSHADER_WARNINGS_AS_ERRORS=OFF -DREACTOR_BACKEND=LLVM
  ninja -C "$BUILD" -j"$(nproc)" vk_swiftshader
fi
cp "$BUILD/Linux/libvk_swiftshader.so" "$OUT/libvk_swiftshader.so"
echo "built $OUT/libvk_swiftshader.so"
