"""Attention backward at the C2 layer-0 shape, a few launches (for ncu captures).  usage: python profiles/exp_attn_bwd_one.py [Lq Lk]"""
import os, sys, torch
sys.path.insert(0, ".")
from recommend_b200 import ops
bf16 = torch.bfloat16
Lq, Lk = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (458, 544)
B, d, H = 2048, 256, 4
g = torch.Generator(device='cuda').manual_seed(0)
rnd = lambda *s: torch.randn(*s, generator=g, device='cuda').to(bf16)
q, kv, do = rnd(Lq * B, d), rnd(Lk * B, 2 * d), rnd(Lq * B, d)
o = torch.empty(Lq * B, d, dtype=bf16, device='cuda'); lse = torch.empty(B * H * Lq, device='cuda')
ops.attn_fwd(q, kv[:, :d], kv[:, d:], o, lse, B, H, Lq, Lk, 64)
dq, dkv = torch.empty_like(q), torch.empty_like(kv); delta = torch.empty(B * H * Lq, device='cuda')
for _ in range(3): ops.attn_bwd(q, kv[:, :d], kv[:, d:], o, lse, do, dq, dkv[:, :d], dkv[:, d:], delta, B, H, Lq, Lk, 64)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): ops.attn_bwd(q, kv[:, :d], kv[:, d:], o, lse, do, dq, dkv[:, :d], dkv[:, d:], delta, B, H, Lq, Lk, 64)
e1.record(); torch.cuda.synchronize()
print(Lq, Lk, f'attn_bwd {e0.elapsed_time(e1)/5:.3f} ms', flush=True)
