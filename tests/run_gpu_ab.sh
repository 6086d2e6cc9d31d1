#!/bin/bash
# Round-2 A/B of the compile-time experiment switches (profiles/README.md, known gaps 1 and 2) in ONE gpurun call:
#   gpurun --timeout 1200 -- 'bash tests/run_gpu_ab.sh'
# Variants: the default build, the FMA-pipe exponential in the attention softmax loops (OT_EX2_POLY_MODE 1 / 2) and the late drain
# wait of the dual-output FFN-1 epilogue (OT_DUAL_LATE_WAIT).  Each variant: rebuild the affected kernels (nvcc on the GPU box,
# same image), run the kernel + model parity tests that exercise them, then a short bench.  The default build is restored at the end.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OBJS="recommend_b200/csrc/obj/ot_attn_fwd_ws.o recommend_b200/csrc/obj/ot_attn_bwd_fused.o recommend_b200/csrc/obj/ot_gemm.o"
i=0
for flags in "" "-DOT_EX2_POLY_MODE=1" "-DOT_EX2_POLY_MODE=2" "-DOT_DUAL_LATE_WAIT=1" ${AB_EXTRA_VARIANTS}; do
  rm -f $OBJS
  OT_NVCC_EXTRA="$flags" bash recommend_b200/csrc/build.sh > gpurun_out/ab_build_$i.log 2>&1 || { echo "[$flags] build failed"; tail -5 gpurun_out/ab_build_$i.log; i=$((i+1)); continue; }
  echo "== variant $i: ${flags:-default}"
  timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -m gpu -x -k "attention or gemm or c1_small or c2_sequence or t9_auc or product_equals or smoke_shape" 2>&1 | tail -2
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
k={x['kernel']:x['ms_per_step'] for x in d.get('kernels',[])}
print('ms/step %.2f  gemm %.2f  attn_fwd %.2f  attn_bwd %.2f' % (d['ms_per_step'], k.get('ot_mixed_gemm',0), k.get('ot_attn_fwd',0), k.get('ot_attn_bwd',0)))" | tee gpurun_out/ab_bench_$i.log
  i=$((i+1))
done
rm -f $OBJS
bash recommend_b200/csrc/build.sh > /dev/null 2>&1 && echo "default build restored"
# run-time switches of the default build (no rebuild): row slices per SM in the weight-gradient kernel
for waves in 3 4; do
  echo "== default build, OT_WGRAD_WAVES=$waves"
  OT_WGRAD_WAVES=$waves timeout 200 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "wgrad" 2>&1 | tail -1
  OT_WGRAD_WAVES=$waves timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
k={x['kernel']:x['ms_per_step'] for x in d.get('kernels',[])}
print('ms/step %.2f  wgrad %.2f' % (d['ms_per_step'], k.get('ot_wgrad',0)))" | tee gpurun_out/ab_bench_wgrad_$waves.log
done
