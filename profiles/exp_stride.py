import sys, torch
sys.path.insert(0, '/root/repo')
from recommend_b200 import ops
bf16 = torch.bfloat16
def run(B, ldpad):
    Lq, Lk, d, H = 458, 544, 256, 4
    g = torch.Generator(device='cuda').manual_seed(0)
    q = torch.randn(Lq * B, d, generator=g, device='cuda').to(bf16)
    kvbuf = torch.randn(Lk * B, 2 * d + ldpad, generator=g, device='cuda').to(bf16)
    kv = kvbuf[:, :2 * d]
    o = torch.empty(Lq * B, d, dtype=bf16, device='cuda'); lse = torch.empty(B * H * Lq, device='cuda')
    for _ in range(2): ops.attn_fwd(q, kv[:, :d], kv[:, d:], o, lse, B, H, Lq, Lk, 64)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): ops.attn_fwd(q, kv[:, :d], kv[:, d:], o, lse, B, H, Lq, Lk, 64)
    e1.record(); torch.cuda.synchronize()
    print(f'attn_fwd B={B} ldpad={ldpad}: {e0.elapsed_time(e1)/5*1e3/B:.3f} us/sample')
for B, pad in [(2048, 0), (2040, 0), (2048, 8), (2048, 64), (2000, 0)]: run(B, pad)
