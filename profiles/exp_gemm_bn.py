"""FFN-1 forward / FFN-2 input-gradient GEMMs at layer-0 size of C2 with block_n 256 vs 128 (with 128 the four epilogue
sets split into two groups that alternate over tiles)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recommend_b200 import ops
from recommend_b200._lib import OT_EPI_BIAS, OT_EPI_GELU, OT_EPI_GELU_GRAD, OT_EPI_RESIDUAL
bf16 = torch.bfloat16
B, Lq, Lk, d, F = 2048, 458, 544, 256, 1024
rows = Lq * B
g = torch.Generator(device='cuda').manual_seed(0)
rnd = lambda *s: (torch.randn(*s, generator=g, device='cuda')).to(bf16)
segs = ops.position_segments(Lk - Lq, Lk, Lk, 32, 'tail', B)
zn, W1 = rnd(rows, d), rnd(33, F, d) * 0.1
b1 = torch.randn(33, F, device='cuda')
h, pre = torch.empty(rows, F, dtype=bf16, device='cuda'), torch.empty(rows, F, dtype=bf16, device='cuda')
dy, W2b = rnd(rows, d), rnd(33, F, d) * 0.1
dpre = torch.empty(rows, F, dtype=bf16, device='cuda')
o, Wo = rnd(rows, d), rnd(1, d, d) * 0.1
z = torch.empty(rows, d, dtype=bf16, device='cuda')


def timed(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

ref = {}
for bn in (256, 128):
    t3 = timed(lambda: ops.mixed_gemm(zn, W1, segs, h, flags=OT_EPI_BIAS | OT_EPI_GELU, bias=b1, out2=pre, block_n=bn))
    t8 = timed(lambda: ops.mixed_gemm(dy, W2b, segs, dpre, flags=OT_EPI_GELU_GRAD, aux=pre, block_n=bn))
    t4 = timed(lambda: ops.mixed_gemm(o, Wo, [(0, 1, rows, 0, 0)], z, flags=OT_EPI_RESIDUAL, res=zn, block_n=bn))
    gb3, gb8, gb4 = rows * (512 + 4096) / 1e3, rows * (512 + 4096) / 1e3, rows * 1536 / 1e3
    print(f'block_n={bn}: FFN-1 fwd {t3:7.1f} us ({gb3 / t3:6.0f} GB/s)   FFN-2 dgrad {t8:7.1f} us ({gb8 / t8:6.0f} GB/s)   Wo+res {t4:7.1f} us ({gb4 / t4:6.0f} GB/s)')
    if bn == 256: ref = dict(h=h.clone(), pre=pre.clone(), dpre=dpre.clone(), z=z.clone())
    else:
        for k, v in dict(h=h, pre=pre, dpre=dpre, z=z).items():
            print('  max |diff| vs block_n=256', k, float((v.float() - ref[k].float()).abs().max()))
