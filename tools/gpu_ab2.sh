#!/bin/sh
# A/B of library builds on config 2 (full bench, device resident) and config 3: tools/gpu_ab2.sh lib1.so lib2.so ...
for L in "$@"; do
  r=$(A52_B200_LIB=$PWD/$L python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e --no-extra 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read());print(round(d['value']))")
  c=$(A52_B200_LIB=$PWD/$L python tools/dev_c3.py 2072 32 stereo 3 2>/dev/null | tail -1)
  echo "$L: c2 $r | $c"
done
