#!/bin/sh
# A/B of library builds on the GPU box: tools/gpu_ab.sh TAG lib1.so lib2.so ...   (full-size bench, device-resident only)
TAG=$1; shift
for L in "$@"; do
  n=$(basename $L .so)
  A52_B200_LIB=$PWD/$L python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e --no-extra > gpurun_out/${TAG}_$n.json 2> gpurun_out/${TAG}_$n.err
  echo "$n: $(python -c "import json;d=json.load(open('gpurun_out/${TAG}_$n.json'));print(round(d['value']), d['roofline']['frac'])" 2>&1 | tail -1)"
done
