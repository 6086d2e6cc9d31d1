#!/bin/bash
# Round-2 A/B of the FMA-pipe exponential in the attention softmax loops (profiles/README.md, known gap 1), ONE gpurun call:
#   gpurun --timeout 900 -- 'bash tests/run_gpu_ab_ex2.sh'
# For OT_EX2_POLY_MODE in 0 1 2: rebuild the two attention kernels with the switch, run the attention + model parity tests, then a
# short bench.  Leaves the default build (mode 0) in place at the end.  nvcc runs on the GPU box (same image), ~40 s per mode.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for mode in 0 1 2; do
  rm -f recommend_b200/csrc/obj/ot_attn_fwd_ws.o recommend_b200/csrc/obj/ot_attn_bwd_fused.o
  OT_NVCC_EXTRA="-DOT_EX2_POLY_MODE=$mode" bash recommend_b200/csrc/build.sh > gpurun_out/ab_ex2_build_$mode.log 2>&1 || { echo "mode $mode: build failed"; tail -5 gpurun_out/ab_ex2_build_$mode.log; continue; }
  echo "== OT_EX2_POLY_MODE=$mode"
  timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -m gpu -x -k "attention or c1_small or c2_sequence or t9_auc or product_equals" 2>&1 | tail -2
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
k={x['kernel']:x['ms_per_step'] for x in d.get('kernels',[])}
print('ms/step %.2f  attn_fwd %.2f  attn_bwd %.2f' % (d['ms_per_step'], k.get('ot_attn_fwd',0), k.get('ot_attn_bwd',0)))" | tee gpurun_out/ab_ex2_bench_$mode.log
done
rm -f recommend_b200/csrc/obj/ot_attn_fwd_ws.o recommend_b200/csrc/obj/ot_attn_bwd_fused.o
bash recommend_b200/csrc/build.sh > /dev/null 2>&1 && echo "default build restored"
