#!/bin/sh
# compute-sanitizer over the parity tests that exercise the kernel's synchronisation (named barriers, the
# cross-CTA slice hand-over, the L2-resident plan, TMA staging with reads past the frame end) and the encoder.
# usage: tools/gpu_sanitize.sh TAG "tool1 tool2 ..."
TAG=${1:-san}; TOOLS=${2:-"memcheck racecheck initcheck synccheck"}
SEL='time_slicing_is_bit_identical or host_pipeline_chunks or frame_independent or golden_vectors or transient_640k or ragged_and_empty or time_slices_are_invisible or carry_split'
for t in $TOOLS; do
  extra=""
  [ "$t" = "racecheck" ] && extra="--racecheck-report analysis"
  timeout 1500 compute-sanitizer --tool $t $extra --print-limit 20 --error-exitcode 0 \
     python -m pytest tests/test_parity_gpu.py tests/test_round2_gpu.py tests/test_encoder_gpu.py -q -x --tb=line -k "$SEL" \
     > gpurun_out/${TAG}_$t.log 2>&1
  echo "== $t: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/${TAG}_$t.log | tr '\n' ' ')"
done
