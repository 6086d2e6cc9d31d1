#!/bin/bash
# one GPU visit: native probes, pytest -m gpu, smoke, a short bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
if [ -n "$PROBES" ]; then tests/native/run_probe.sh tests/native/bin/probe_gemm $PROBES 2>&1 | grep -E "FAIL|exit code|PASS" | tail -30; fi
OT_LOGIT_TOL=${OT_LOGIT_TOL:-1e-2} OT_GRAD_TOL=${OT_GRAD_TOL:-3e-2} timeout 900 python -m pytest tests -q -m gpu -s ${PYTEST_ARGS} 2>&1 | grep -vE "^\s*$" | grep -E "logits rel-L2|passed|failed|Error|error|FAILED|assert" | tail -${TAILN:-40}
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --steps ${STEPS:-5} --warmup 3 ${BENCH_ARGS} 2>&1 | tail -5 | tee gpurun_out/bench_line.log | cut -c1-600
