"""Smallest invocation of every hand-rolled-synchronisation kernel (mbarrier rings, named barriers, dynamic tile counters) for
compute-sanitizer: grouped GEMM with fused epilogues, fused FFN, weight gradient (plain and with the GELU transform), both attention
forward kernels, the fused attention backward, cached attention.  Run by tests/run_sanitizer.sh; one tool per GPU visit."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recommend_b200 import ops
from recommend_b200._lib import OT_EPI_BIAS, OT_EPI_GELU, OT_EPI_RESIDUAL

bf16 = torch.bfloat16
g = torch.Generator(device='cuda').manual_seed(0)
rnd = lambda *s: torch.randn(*s, generator=g, device='cuda').to(bf16)
B, d, F, H = 4, 256, 256, 4
n_s, n_ns = 40, 2
rows = (n_s + n_ns) * B
segs = [(0, 1, n_s * B, 0, 0), (n_s * B, n_ns, B, 1, 1)]
x, W1, W2 = rnd(rows, d), rnd(3, F, d) * 0.1, rnd(3, d, F) * 0.1
b1, b2 = torch.randn(3, F, device='cuda'), torch.randn(3, d, device='cuda')
gain = torch.ones(d, device='cuda')
h, pre = torch.empty(rows, F, dtype=bf16, device='cuda'), torch.empty(rows, F, dtype=bf16, device='cuda')
y, yn = torch.empty(rows, d, dtype=bf16, device='cuda'), torch.empty(rows, d, dtype=bf16, device='cuda')
rstd = torch.empty(rows, device='cuda')
ops.mixed_gemm(x, W1, segs, h, flags=OT_EPI_BIAS | OT_EPI_GELU, bias=b1, out2=pre)
ops.mixed_gemm(h, W2, segs, y, flags=OT_EPI_BIAS | OT_EPI_RESIDUAL, bias=b2, res=x, dropout=(3, 0.1), norm=(yn, gain, rstd, 1e-6))
ops.ffn_fused(x, W1, b1, W2, b2, segs, y, pre=pre, res=x, dropout=(3, 0.1), norm=(yn, gain, rstd, 1e-6))
dW = torch.zeros(3, F, d, device='cuda')
db = torch.zeros(3, d, device='cuda')
ops.wgrad_rows(h, y, segs, dW, F * d, d, 1, q_colsum=db, q_colsum_group_stride=d)
ops.wgrad_rows(pre, y, segs, dW, F * d, d, 1, q_colsum=db, q_colsum_group_stride=d, p_gelu=True)
for (Lq, Lk) in ((150, 170), (20, 42)):            # two query tiles (v2 forward) and a short tail (round-1 forward)
    q, kv, do = rnd(Lq * B, d), rnd(Lk * B, 2 * d), rnd(Lq * B, d)
    o = torch.empty(Lq * B, d, dtype=bf16, device='cuda')
    lse = torch.empty(B * H * Lq, device='cuda')
    ops.attn_fwd(q, kv[:, :d], kv[:, d:], o, lse, B, H, Lq, Lk, d // H)
    dq, dkv = torch.empty_like(q), torch.empty_like(kv)
    delta = torch.empty(B * H * Lq, device='cuda')
    ops.attn_bwd(q, kv[:, :d], kv[:, d:], o, lse, do, dq, dkv[:, :d], dkv[:, d:], delta, B, H, Lq, Lk, d // H)
C, Tq, Tn, Ls = 8, 4, 4, 100
q, kvo, kvs = rnd(Tq * C, d), rnd(Tn * C, 2 * d), rnd(Ls, 2 * d)
o = torch.empty(Tq * C, d, dtype=bf16, device='cuda')
ops.attn_ns_cached(q, kvo[:, :d], kvo[:, d:], kvs[:, :d], kvs[:, d:], o, C, H, Tq, Tn, Ls, d // H)
torch.cuda.synchronize()
print('sanitizer smoke: all launches completed')
