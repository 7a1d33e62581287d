#!/bin/sh
# Builds liba52_b200.so (CUDA kernels + C ABI + drop-in liba52 API) for sm_100a, in tree.
set -e
here=$(cd "$(dirname "$0")" && pwd)
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
    -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -shared \
    -Xptxas -v -Xlinker --version-script="$here/csrc/exports.map" \
    -o "$here/liba52_b200.so" "$here/csrc/a52_decode.cu" "$here/csrc/ac3_encode.cu" -lcudart "$@"
# host tools over the C ABI (C, no CUDA in them): the a52dec command line on the batched engine
CC=${CC:-gcc}
$CC -O2 -std=gnu99 -Wall -Wextra -o "$here/a52dec_b200" "$here/cli/a52dec_b200.c" \
    -I"$here/../include" -L"$here" -l:liba52_b200.so -Wl,-rpath,'$ORIGIN' -lm
$CC -O2 -std=gnu99 -Wall -Wextra -o "$here/ac3enc_b200" "$here/cli/ac3enc_b200.c" \
    -I"$here/../include" -L"$here" -l:liba52_b200.so -Wl,-rpath,'$ORIGIN' -lm
