#!/bin/bash
# one GPU visit: native probes, pytest -m gpu, smoke, a short bench.  Stops at the first hang.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
if [ -n "$PROBES" ]; then
  tests/native/run_probe.sh tests/native/bin/probe_gemm $PROBES > gpurun_out/probes.log 2>&1; prc=$?
  grep -E "FAIL|exit code|PASS|hang" gpurun_out/probes.log | tail -30
  if [ $prc -eq 124 ]; then echo "probe hang -> skipping the rest of the round"; exit 124; fi
fi
OT_LOGIT_TOL=${OT_LOGIT_TOL:-1e-2} OT_GRAD_TOL=${OT_GRAD_TOL:-3e-2} timeout ${PYTEST_TIMEOUT:-600} python -m pytest tests -q -m gpu -s -x --timeout 120 ${PYTEST_ARGS} 2>&1 | grep -vE "^\s*$" | grep -E "logits rel-L2|passed|failed|Error|error|FAILED|assert|Timeout" | tail -${TAILN:-40}
if [ ${PIPESTATUS[0]} -eq 124 ]; then echo "pytest hang -> skipping the rest of the round"; exit 124; fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --steps ${STEPS:-5} --warmup 3 ${BENCH_ARGS} 2>&1 | tail -5 | tee gpurun_out/bench_line.log | cut -c1-600
