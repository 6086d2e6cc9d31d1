#!/bin/bash
# quick GPU bring-up: model-level parity tests with full output
cd "$(dirname "$0")/.."
OT_LOGIT_TOL=${OT_LOGIT_TOL:-1e-2} OT_GRAD_TOL=${OT_GRAD_TOL:-3e-2} python -m pytest tests/test_gpu_model.py -q -m gpu -s 2>&1 | grep -vE "^\s*$" | tail -${TAILN:-80}
