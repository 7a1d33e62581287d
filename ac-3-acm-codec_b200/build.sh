#!/bin/sh
# Builds liba52_b200.so (CUDA kernels + C ABI + drop-in liba52 API) for sm_100a, in tree.
set -e
here=$(cd "$(dirname "$0")" && pwd)
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
    -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -shared \
    -Xptxas -v -Xlinker --version-script="$here/csrc/exports.map" \
    -o "$here/liba52_b200.so" "$here/csrc/a52_decode.cu" "$here/csrc/ac3_encode.cu" -lcudart "$@"
