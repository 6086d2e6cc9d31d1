#!/bin/bash
# compute-sanitizer over the smallest launch of every kernel with hand-rolled synchronisation.
#   gpurun --timeout 900 -- 'bash tests/run_sanitizer.sh synccheck'      (one tool per GPU visit: B200_PROFILING.md)
# The plain run must exit 0 first; the log lands in gpurun_out/ and its summary is copied to profiles/ by hand.
cd "$(dirname "$0")/.."
TOOL=${1:-synccheck}
mkdir -p gpurun_out
timeout 120 python tests/sanitizer_smoke.py > gpurun_out/sanitizer_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/sanitizer_plain.log; exit 1; }
timeout ${SAN_TIMEOUT:-600} compute-sanitizer --tool "$TOOL" --print-limit 20 python tests/sanitizer_smoke.py > gpurun_out/sanitizer_${TOOL}.log 2>&1
echo "compute-sanitizer --tool $TOOL exit code $?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitizer smoke|hazard|Barrier error|Invalid" gpurun_out/sanitizer_${TOOL}.log | head -20
