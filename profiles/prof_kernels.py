"""Launches the hot kernels once each at BASELINE config 2 / layer 0 sizes (B 2048, 458 query rows of 544) so that
ncu can capture them in isolation:  python profiles/prof_kernels.py [gemm|ffn|attn|r2|all]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recommend_b200 import ops
from recommend_b200._lib import OT_EPI_BIAS, OT_EPI_GELU, OT_EPI_GELU_GRAD, OT_EPI_RESIDUAL

what = sys.argv[1] if len(sys.argv) > 1 else 'all'
REP = 1 if os.environ.get('PROF_ONCE') else 2      # launches per kernel (the first one is the cold-cache one)
prof = ops.KernelProfiler()      # records (family, shape tag, algorithmic flops / bytes) of every launch, in order
ops.set_profiler(prof)
bf16 = torch.bfloat16
B, Lq, Lk, d, F, H = 2048, 458, 544, 256, 1024, 4
rows = Lq * B
g = torch.Generator(device='cuda').manual_seed(0)
rnd = lambda *s: (torch.randn(*s, generator=g, device='cuda')).to(bf16)
segs = ops.position_segments(Lk - Lq, Lk, Lk, 32, 'tail', B)
if what in ('gemm', 'all'):
    zn, W1 = rnd(rows, d), rnd(33, F, d) * 0.1
    b1 = torch.randn(33, F, device='cuda')
    h, pre = torch.empty(rows, F, dtype=bf16, device='cuda'), torch.empty(rows, F, dtype=bf16, device='cuda')
    for _ in range(REP):
        ops.mixed_gemm(zn, W1, segs, h, flags=OT_EPI_BIAS | OT_EPI_GELU, bias=b1, out2=pre)          # FFN-1 forward
    dy, W2b = rnd(rows, d), rnd(33, F, d) * 0.1
    dpre = torch.empty(rows, F, dtype=bf16, device='cuda')
    for _ in range(REP):
        ops.mixed_gemm(dy, W2b, segs, dpre, flags=OT_EPI_GELU_GRAD, aux=pre)                          # FFN-2 input gradient
    W2 = rnd(33, d, F) * 0.1
    y = torch.empty(rows, d, dtype=bf16, device='cuda')
    for _ in range(REP):
        ops.mixed_gemm(h, W2, segs, y, flags=OT_EPI_BIAS | OT_EPI_RESIDUAL, bias=b1[:, :d].contiguous(), res=zn)   # FFN-2 forward
    dW = torch.zeros(33, d, F, device='cuda')
    for _ in range(REP):
        ops.wgrad_rows(zn, dpre, segs, dW, d * F, F, 1)                                               # dW1
if what in ('ffn', 'all'):
    # round 2: fused FFN forward (h stays on chip) and dW2 with the GELU rebuilt inside the weight-gradient kernel
    zn, W1, W2 = rnd(rows, d), rnd(33, F, d) * 0.06, rnd(33, d, F) * 0.03
    b1, b2 = 0.1 * torch.randn(33, F, device='cuda'), 0.1 * torch.randn(33, d, device='cuda')
    res, gain = rnd(rows, d), torch.ones(d, device='cuda')
    y, nout = torch.empty(rows, d, dtype=bf16, device='cuda'), torch.empty(rows, d, dtype=bf16, device='cuda')
    rstd, pre = torch.empty(rows, device='cuda'), torch.empty(rows, F, dtype=bf16, device='cuda')
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(8)]
    for i in range(3 if REP > 1 else 1):
        ev[2 * i].record()
        ops.ffn_fused(zn, W1, b1, W2, b2, segs, y, pre=pre, res=res, dropout=(7, 0.1), norm=(nout, gain, rstd, 1e-6))
        ev[2 * i + 1].record()
    ops.ffn_fused(zn, W1, b1, W2, b2, segs, y, res=res)                                                  # evaluation form: no pre, no norm
    dy = rnd(rows, d)
    dW2, db2 = torch.zeros(33, F, d, device='cuda'), torch.zeros(33, d, device='cuda')
    ev[6].record()
    ops.wgrad_rows(pre, dy, segs, dW2, F * d, d, 1, q_colsum=db2, q_colsum_group_stride=d, p_gelu=True)
    ev[7].record()
    dpre_f, dzn_f = torch.empty(rows, F, dtype=bf16, device='cuda'), torch.empty(rows, d, dtype=bf16, device='cuda')
    ops.ffn_fused_bwd(dy, rnd(33, F, d) * 0.05, rnd(33, d, F) * 0.05, pre, segs, dpre_f, dzn_f)      # fused input-gradient chain
    torch.cuda.synchronize()
    print('ffn_fused layer-0 size (ms):', [round(ev[2 * i].elapsed_time(ev[2 * i + 1]), 3) for i in range(3 if REP > 1 else 1)], 'wgrad gelu:', round(ev[6].elapsed_time(ev[7]), 3))
if what in ('attn', 'all'):
    q, kv, do = rnd(rows, d), rnd(Lk * B, 2 * d), rnd(rows, d)
    o = torch.empty(rows, d, dtype=bf16, device='cuda')
    lse = torch.empty(B * H * Lq, device='cuda')
    for _ in range(REP):
        ops.attn_fwd(q, kv[:, :d], kv[:, d:], o, lse, B, H, Lq, Lk, d // H)
    dq, dkv = torch.empty_like(q), torch.empty_like(kv)
    delta = torch.empty(B * H * Lq, device='cuda')
    for _ in range(REP):
        ops.attn_bwd(q, kv[:, :d], kv[:, d:], o, lse, do, dq, dkv[:, :d], dkv[:, d:], delta, B, H, Lq, Lk, d // H)
if what in ('r2', 'all'):
    # round 2: the remaining buckets >= 1 ms/step the round-1 verdict asked ncu rows for
    from recommend_b200._lib import OT_EPI_BIAS as _B
    o_, Wo = rnd(rows, d), rnd(1, d, d) * 0.06
    res, gain = rnd(rows, d), torch.ones(d, device='cuda')
    z, zn = torch.empty(rows, d, dtype=bf16, device='cuda'), torch.empty(rows, d, dtype=bf16, device='cuda')
    r2 = torch.empty(rows, device='cuda')
    for _ in range(REP):      # Wo + residual + dropout + fused norm2 (OT/model.py:117,193,196)
        ops.mixed_gemm(o_, Wo, [(0, 1, rows, 0, 0)], z, flags=OT_EPI_RESIDUAL, res=res, dropout=(5, 0.1), norm=(zn, gain, r2, 1e-6))
    dzn, dz, dz_a, dy = rnd(rows, d), torch.empty(rows, d, dtype=bf16, device='cuda'), torch.empty(rows, d, dtype=bf16, device='cuda'), rnd(rows, d)
    dg = torch.zeros(d, device='cuda')
    for _ in range(REP):      # norm2 backward with the masked second output
        ops.rmsnorm_bwd(dzn, z, r2, gain, dz, dg, dres=dy, drop_out=(dz_a, 5, 0.1, 0))
    ev, Ws, bs = rnd(B, 170, 64), rnd(1, d, 64) * 0.1, torch.zeros(d, device='cuda')
    X0 = torch.empty(544 * B, d, dtype=bf16, device='cuda')
    for _ in range(REP):      # sequence tokenizer projection, K = 64 (OT/model.py:262-265)
        ops.mixed_gemm(ev, Ws, [(0, 170, B, 0, 0)], X0, flags=_B, bias=bs, a_transposed_events=True)
    C_, Tq, Tn, Ls = 8192, 32, 32, 512
    qc, kvo, kvs = rnd(Tq * C_, d), rnd(Tn * C_, 2 * d), rnd(Ls, 2 * d)
    oc = torch.empty(Tq * C_, d, dtype=bf16, device='cuda')
    for _ in range(REP):      # cached-candidate attention, BASELINE config 5 layer 0
        ops.attn_ns_cached(qc, kvo[:, :d], kvo[:, d:], kvs[:, :d], kvs[:, d:], oc, C_, H, Tq, Tn, Ls, d // H)
    dL, HL, LqL, LkL = 384, 4, 480, 544      # OneTrans-L layer 0: head_dim 96
    ql, kvl, dol = rnd(LqL * B, dL), rnd(LkL * B, 2 * dL), rnd(LqL * B, dL)
    ol = torch.empty(LqL * B, dL, dtype=bf16, device='cuda')
    lsel, dell = torch.empty(B * HL * LqL, device='cuda'), torch.empty(B * HL * LqL, device='cuda')
    ops.attn_fwd(ql, kvl[:, :dL], kvl[:, dL:], ol, lsel, B, HL, LqL, LkL, dL // HL)
    dql, dkvl = torch.empty_like(ql), torch.empty_like(kvl)
    ops.attn_bwd(ql, kvl[:, :dL], kvl[:, dL:], ol, lsel, dol, dql, dkvl[:, :dL], dkvl[:, dL:], dell, B, HL, LqL, LkL, dL // HL)
torch.cuda.synchronize()
import json
json.dump([{'kernel': r[0], 'tag': r[1], 'flops': r[4], 'bytes': r[5]} for r in prof.records],
          open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out', f'prof_kernels_{what}.json'), 'w'), indent=1)
print('done')
