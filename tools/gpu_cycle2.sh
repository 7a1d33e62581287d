#!/bin/sh
# developer loop on the GPU box (round 2): parity tests, the full bench line, a balanced small bench and one ncu capture
# usage: tools/gpu_cycle2.sh TAG [notests]
TAG=${1:-x}
mkdir -p gpurun_out
cp ac-3-acm-codec_b200/csrc/a52_decode.cu gpurun_out/${TAG}_a52_decode.cu
if [ "$2" != "notests" ]; then
  python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -15 > gpurun_out/${TAG}_tests.log
  tail -3 gpurun_out/${TAG}_tests.log
fi
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cut -c1-300 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
export A52_B200_SLICE_FRAMES=64
CMD="python bench.py --streams 1776 --frames 64 --steps 2 --warmup 3 --no-e2e --no-cpu --no-extra"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:a52_decode -s 3 -c 1 -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu.log 2>&1
cut -c1-200 gpurun_out/${TAG}_plain.log; tail -1 gpurun_out/${TAG}_ncu.log
