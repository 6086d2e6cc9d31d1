import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recommend_b200 import ops
bf16 = torch.bfloat16
H, d = 4, 256
for (B, Lq, Lk) in [(2048, 458, 544), (2048, 288, 373), (256, 1024, 2048)]:
    g = torch.Generator(device='cuda').manual_seed(0)
    rnd = lambda *s: torch.randn(*s, generator=g, device='cuda').to(bf16)
    q, kv = rnd(Lq * B, d), rnd(Lk * B, 2 * d)
    o = torch.empty(Lq * B, d, dtype=bf16, device='cuda'); lse = torch.empty(B * H * Lq, device='cuda')
    for sw in (0, 128):
        for _ in range(3): ops.attn_fwd(q, kv[:, :d], kv[:, d:], o, lse, B, H, Lq, Lk, 64, swizzle=sw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): ops.attn_fwd(q, kv[:, :d], kv[:, d:], o, lse, B, H, Lq, Lk, 64, swizzle=sw)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        pairs = Lq * Lk - Lq * (Lq - 1) / 2
        print(f'attn_fwd {"ws" if sw == 0 else "simple"} B={B} Lq={Lq} Lk={Lk}: {ms:.3f} ms  {4 * B * H * 64 * pairs / ms / 1e9:.0f} TFLOP/s')
