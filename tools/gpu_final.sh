#!/bin/sh
# round-end record on one B200: GPU suite, the bench lines (decode, encode, reference arm), the launch list of a bench
# run and one full ncu capture of each kernel.  usage: tools/gpu_final.sh TAG
TAG=${1:-final}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -4 > $O/${TAG}_tests.log; tail -2 $O/${TAG}_tests.log
python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; cut -c1-160 $O/${TAG}_bench.json
python bench.py --workload encode > $O/${TAG}_bench_encode.json 2> $O/${TAG}_bench_encode.err; cut -c1-160 $O/${TAG}_bench_encode.json
python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err; cut -c1-200 $O/${TAG}_bench_reference.json
# launch list of a short bench run (only after the same command ran clean without ncu)
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-extra --e2e-steps 1"
$CMD > $O/${TAG}_l_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv $CMD > $O/${TAG}_l_ncu.log 2>&1
# full captures
cp ac-3-acm-codec_b200/csrc/a52_decode.cu $O/${TAG}_a52_decode.cu
cp ac-3-acm-codec_b200/csrc/ac3_encode.cu $O/${TAG}_ac3_encode.cu
export A52_B200_SLICE_FRAMES=64
CMD="python bench.py --streams 1776 --frames 64 --steps 2 --warmup 3 --no-e2e --no-cpu --no-extra"
$CMD > $O/${TAG}_d_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:a52_decode -s 3 -c 1 -o $O/${TAG}_dec_prof $CMD > $O/${TAG}_d_ncu.log 2>&1
unset A52_B200_SLICE_FRAMES
CMD="python bench.py --workload encode --streams 444 --frames 32 --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > $O/${TAG}_e_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ac3_encode -s 3 -c 1 -o $O/${TAG}_enc_prof $CMD > $O/${TAG}_e_ncu.log 2>&1
tail -1 $O/${TAG}_d_ncu.log; tail -1 $O/${TAG}_e_ncu.log
