#!/bin/sh
# developer loop on the GPU box: parity tests, a balanced small bench, one ncu capture of the decode kernel
python -m pytest tests -m gpu -q --tb=line -x 2>&1 | tail -6
CMD="python bench.py --streams 1776 --frames 32 --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/v2_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:a52_decode -s 3 -c 1 -o gpurun_out/v2_prof $CMD > gpurun_out/v2_ncu.log 2>&1
cut -c1-220 gpurun_out/v2_plain.log; tail -1 gpurun_out/v2_ncu.log
