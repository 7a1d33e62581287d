#!/bin/sh
# A/B of environment settings on the GPU box: tools/gpu_env_ab.sh TAG "VAR=a" "VAR=b" ...   (bench device-resident + config 3)
TAG=$1; shift
for E in "$@"; do
  r=$(env $E python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e --no-extra 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read());print(round(d['value']))")
  c=$(env $E python tools/dev_c3.py 2072 32 stereo 3 2>/dev/null | tail -1)
  echo "$E: c2 $r | $c"
done
