#!/bin/bash
# Runs every probe test in its own process (a device fault in one cannot poison the next).
# usage: run_probe.sh <binary> <ids...>
BIN=$1; shift
fail=0
for id in "$@"; do
  timeout 120 "$BIN" "$id"
  rc=$?
  if [ $rc -ne 0 ]; then echo "  -> test $id exit code $rc"; fail=1; fi
done
exit $fail
