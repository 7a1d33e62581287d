#!/bin/sh
# per-stage breakdown of an encoder capture: tools/ncu_enc_lines.sh TAG [frames]
TAG=$1
mkdir -p /tmp/$TAG && cp gpurun_out/${TAG}_ac3_encode.cu /tmp/$TAG/ac3_encode.cu
ncu -i gpurun_out/${TAG}_prof.ncu-rep --page source --csv --print-source cuda,sass > /tmp/$TAG/src.csv 2>/dev/null
python tools/ncu_lines.py /tmp/$TAG/src.csv /tmp/$TAG/ac3_encode.cu ${2:-14208}
