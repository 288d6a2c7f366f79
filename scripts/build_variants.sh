#!/bin/bash
# Builds differently tuned copies of libswcuda.so into gpurun_out/variants/ (they travel to the GPU box with the snapshot? no:
# gpurun_out/ is not sent) -> build into swiftshader_b200/csrc/variants/ (git-ignored *.so).  usage: scripts/build_variants.sh name "EXTRA flags" ...
set -e
cd "$(dirname "$0")/../swiftshader_b200/csrc"
mkdir -p variants
while [ $# -ge 2 ]; do
  name="$1"; flags="$2"; shift 2
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a $flags -O3 -std=c++17 -lineinfo -fmad=false -ftz=true -prec-div=true -prec-sqrt=true \
    -Xcompiler -fPIC,-ffp-contract=off -shared -o variants/libswcuda_$name.so draw.cu spirv_subset.cpp -lcudart &
done
wait
ls -la variants
