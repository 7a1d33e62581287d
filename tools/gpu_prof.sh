#!/bin/sh
# one ncu capture of the decode kernel for a given library build: tools/gpu_prof.sh TAG lib.so [streams]
TAG=$1; L=$2; S=${3:-1776}
cp ac-3-acm-codec_b200/csrc/a52_decode.cu gpurun_out/${TAG}_a52_decode.cu
export A52_B200_LIB=$PWD/$L A52_B200_SLICE_FRAMES=64
CMD="python bench.py --streams $S --frames 64 --steps 2 --warmup 3 --no-e2e --no-cpu --no-extra"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:a52_decode -s 3 -c 1 -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu.log 2>&1
cut -c1-200 gpurun_out/${TAG}_plain.log; tail -1 gpurun_out/${TAG}_ncu.log
