#!/bin/sh
# instruction count of one kernel in the built library: tools/sass_size.sh [pattern]
cuobjdump -sass ac-3-acm-codec_b200/liba52_b200.so | awk -v pat="${1:-ac3_encode_kernel}" '
  /Function :/ { f = index($0, pat) > 0 }
  f && /^ +\/\*[0-9a-f]+\*\/ +[A-Z@]/ { n++ }
  END { print n " instructions, " n * 16 / 1024 " KB" }'
