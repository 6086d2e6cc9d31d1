#!/bin/bash
# Runs every probe test in its own process (a device fault in one cannot poison the next).
# A probe that times out (hung kernel) stops the run: the remaining ones would only burn GPU time.
# usage: run_probe.sh <binary> <ids...>
BIN=$1; shift
fail=0
for id in "$@"; do
  timeout ${PROBE_TIMEOUT:-40} "$BIN" "$id"
  rc=$?
  if [ $rc -ne 0 ]; then echo "  -> test $id exit code $rc"; fail=1; fi
  if [ $rc -eq 124 ]; then echo "  -> hang: aborting remaining probes"; exit 124; fi
done
exit $fail
