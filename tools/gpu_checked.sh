#!/bin/sh
# The whole GPU suite against a bounds-checked build of the decode kernel (-DA52_BOUNDS_CHECK: shared-memory store
# addresses, plan indices, bit-window words, staged-frame sizes), compute-sanitizer being closed on this pool.
# usage: tools/gpu_checked.sh TAG     (run on the GPU box; builds with nvcc there)
TAG=${1:-chk}
mkdir -p build gpurun_out
here=ac-3-acm-codec_b200
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
  -shared -Xlinker --version-script=$here/csrc/exports.map -DA52_BOUNDS_CHECK -DA52_PAIRS_PER_CTA=12 -o build/liba52_checked.so \
  $here/csrc/a52_decode.cu $here/csrc/ac3_encode.cu -lcudart || exit 1
A52_B200_LIB=$PWD/build/liba52_checked.so python -m pytest tests -m gpu -q --tb=short -s 2>&1 | tail -12 | tee gpurun_out/${TAG}_checked.log
