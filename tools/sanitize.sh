#!/bin/bash
# compute-sanitizer over the frame path (run under gpurun): memcheck (global / shared out-of-bounds, misaligned,
# leaks of the API) and racecheck (shared-memory hazards; the relabelling pass's "only this thread writes it"
# ownership argument is about global memory and is argued in ssf_tps.cu, racecheck covers the staged tiles and the
# block reductions).  Logs land in gpurun_out/sanitizer_<tool>_<tag>.log; copy the summaries to profiles/.
TAG=${1:-rX}
OUT=gpurun_out
mkdir -p $OUT
for TOOL in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $TOOL --print-limit 20 --log-file $OUT/sanitizer_${TOOL}_$TAG.log \
    python tools/sanitize_frame.py > $OUT/sanitizer_${TOOL}_$TAG.out 2>&1
  echo "$TOOL exit $?"; tail -3 $OUT/sanitizer_${TOOL}_$TAG.out; tail -5 $OUT/sanitizer_${TOOL}_$TAG.log
done
