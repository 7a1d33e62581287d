#!/bin/sh
# one ncu capture of the encode kernel: tools/gpu_prof_enc.sh TAG
TAG=$1
cp ac-3-acm-codec_b200/csrc/ac3_encode.cu gpurun_out/${TAG}_ac3_encode.cu
CMD="python bench.py --workload encode --streams 444 --frames 32 --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ac3_encode -s 3 -c 1 -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu.log 2>&1
cut -c1-200 gpurun_out/${TAG}_plain.log; tail -1 gpurun_out/${TAG}_ncu.log
python bench.py --workload encode --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | cut -c1-160
