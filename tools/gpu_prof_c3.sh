#!/bin/sh
TAG=$1; MODE=${2:-stereo}
cp ac-3-acm-codec_b200/csrc/a52_decode.cu gpurun_out/${TAG}_a52_decode.cu
CMD="python tools/dev_c3.py 2072 32 $MODE 2"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:a52_decode -s 3 -c 1 -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu.log 2>&1
cat gpurun_out/${TAG}_plain.log | tail -2; tail -1 gpurun_out/${TAG}_ncu.log
