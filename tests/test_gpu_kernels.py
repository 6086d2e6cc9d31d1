"""GPU parity of every kernel behind the C-ABI (called through recommend_b200.ops -> ctypes -> libonetrans_sm100.so)
against plain PyTorch fp32 references of the same op on the same seeded inputs, including edge cases (ragged
tiles, partial groups, tails shorter than a tile) and BASELINE.json's full sizes."""
import os
import math

import pytest
import torch

from recommend_b200 import ops, _lib
from recommend_b200._lib import OT_EPI_BIAS, OT_EPI_GELU, OT_EPI_GELU_GRAD, OT_EPI_RESIDUAL, OT_EPI_ROW_SCALE

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16
TOL = 2e-2   # bf16 outputs vs fp32 reference: |got-ref| <= TOL * (1 + |ref|)


def close(got, ref, tol=TOL):
    got, ref = got.float(), ref.float()
    err = ((got - ref).abs() / (1 + ref.abs())).max().item()
    assert err <= tol, f'max scaled err {err:.3e}'


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator(device='cuda').manual_seed(seed)
    return (torch.randn(*shape, generator=g, device='cuda') * scale).to(bf16)


def ref_rows_gemm(A, W, segs):
    out = torch.zeros(A.shape[0], W.shape[1], device='cuda')
    for (r0, n_units, rpu, g0, gs) in segs:
        for u in range(n_units):
            sl = slice(r0 + u * rpu, r0 + (u + 1) * rpu)
            out[sl] = A[sl].float() @ W[g0 + u * gs].float().t()
    return out


@pytest.mark.parametrize('rows,N,K,bn', [(128, 64, 64, 0), (1000, 256, 256, 256), (777, 768, 256, 0), (300, 384, 384, 128),
                                         (4096 + 13, 1536, 384, 0)])
def test_gemm_plain(rows, N, K, bn):
    A, W = rnd(rows, K, seed=1), rnd(1, N, K, scale=0.1, seed=2)
    out = torch.empty(rows, N, dtype=bf16, device='cuda')
    ops.mixed_gemm(A, W, [(0, 1, rows, 0, 0)], out, block_n=bn)
    close(out, A.float() @ W[0].float().t())


@pytest.mark.parametrize('B', [32, 128, 200])
def test_gemm_grouped_mixed_parameters(B):
    """Shared S run + per-token NS runs in one launch (OT/model.py:84-92), partial tiles when B % 128 != 0."""
    n_s, n_ns, K, N = 5, 6, 256, 512
    rows = (n_s + n_ns) * B
    A, W = rnd(rows, K, seed=3), rnd(1 + 8, N, K, scale=0.1, seed=4)
    segs = [(0, 1, n_s * B, 0, 0), (n_s * B, n_ns, B, 3, 1)]     # NS tokens 2..7 -> groups 3..8
    out = torch.full((rows, N), float('nan'), dtype=bf16, device='cuda')
    ops.mixed_gemm(A, W, segs, out)
    close(out, ref_rows_gemm(A, W, segs))


def test_gemm_epilogues():
    rows, N, K, G = 600, 1024, 256, 3
    A, W = rnd(rows, K, seed=5), rnd(G, N, K, scale=0.1, seed=6)
    bias = torch.randn(G, N, device='cuda')
    res, aux = rnd(rows, N, seed=7), rnd(rows, N, seed=8)
    rs = torch.rand(rows, device='cuda') + 0.5
    segs = [(0, 1, 344, 0, 0), (344, 2, 128, 1, 1)]
    base = ref_rows_gemm(A, W, segs)
    gidx = torch.cat([torch.zeros(344), torch.ones(128), 2 * torch.ones(128)]).long().cuda()
    # bias + gelu with the pre-activation kept
    out, pre = torch.empty(rows, N, dtype=bf16, device='cuda'), torch.empty(rows, N, dtype=bf16, device='cuda')
    ops.mixed_gemm(A, W, segs, out, flags=OT_EPI_BIAS | OT_EPI_GELU, bias=bias, out2=pre)
    close(pre, base + bias[gidx])
    close(out, torch.nn.functional.gelu(base + bias[gidx]))
    # bias + residual
    ops.mixed_gemm(A, W, segs, out, flags=OT_EPI_BIAS | OT_EPI_RESIDUAL, bias=bias, res=res)
    close(out, base + bias[gidx] + res.float())
    # gelu' multiply (backward of the FFN activation), row scale
    x = aux.float().requires_grad_(True)
    torch.nn.functional.gelu(x).sum().backward()
    ops.mixed_gemm(A, W, segs, out, flags=OT_EPI_GELU_GRAD | OT_EPI_ROW_SCALE, aux=aux, row_scale=rs)
    close(out, base * rs[:, None] * x.grad)
    # in-place residual accumulation (out aliases res), as the Q-projection input gradient uses it
    acc = res.clone()
    ops.mixed_gemm(A, W, segs, acc, flags=OT_EPI_RESIDUAL, res=acc)
    close(acc, base + res.float())


@pytest.mark.parametrize('rows,N,K,hp_rows', [(1000, 256, 256, 0), (1280, 256, 1024, 256), (300, 128, 64, 0), (77, 64, 128, 0)])
def test_gemm_fused_rmsnorm_output(rows, N, K, hp_rows):
    """OT_EPI_NORM: the epilogue that writes z = res + A W^T also writes RMSNorm(z) * gain and rstd (OT/model.py:19-23)."""
    A, W = rnd(rows, K, seed=31), rnd(1, N, K, seed=32) * 0.1
    res = rnd(rows, N, seed=33)
    gain = (1.0 + 0.1 * torch.randn(N, device='cuda'))
    out = torch.empty(rows, N, dtype=bf16, device='cuda')
    nout = torch.empty(rows, N, dtype=bf16, device='cuda')
    rstd = torch.empty(rows, device='cuda')
    kw = {}
    segs = [(0, 1, rows, 0, 0)]
    if hp_rows:
        res_hp = res[rows - hp_rows:].float() + 1e-3 * torch.randn(hp_rows, N, device='cuda')
        out_hp = torch.empty_like(res_hp)
        kw = dict(res_hp=res_hp, out_hp=out_hp, hp_row0=rows - hp_rows)
        segs = [(0, 1, rows - hp_rows, 0, 0), (rows - hp_rows, 1, hp_rows, 0, 0)]
    ops.mixed_gemm(A, W, segs, out, flags=OT_EPI_RESIDUAL, res=res, norm=(nout, gain, rstd, 1e-6), **kw)
    z = A.float() @ W[0].float().t() + res.float()
    if hp_rows:
        z[rows - hp_rows:] = A[rows - hp_rows:].float() @ W[0].float().t() + res_hp
        assert (out_hp - z[rows - hp_rows:]).abs().max().item() < 2e-3 * z.abs().max().item()
    assert ((out.float() - z).abs().max() / z.abs().max()).item() < 1e-2
    # row statistics come from the fp32 results (before the bf16 rounding of `out`), the scaled values from what went to HBM
    zq = out.float()
    if hp_rows:
        zq[rows - hp_rows:] = out_hp
    r_ref = torch.rsqrt(z.square().mean(-1) + 1e-6)
    assert ((rstd - r_ref).abs().max() / r_ref.abs().max()).item() < 2e-4      # z itself is a bf16-input matmul: fp32 accumulation order
    n_ref = zq * rstd[:, None] * gain
    assert ((nout.float() - n_ref).abs().max() / n_ref.abs().max()).item() < 6e-3      # one bf16 rounding


def test_gemm_transposed_events_tokenizer_layout():
    """[B, L_i, 64] events -> token-major rows (l, b) (OT/model.py:262-265 + DESIGN.md layout)."""
    B, Li, E, d, off = 200, 7, 64, 256, 3
    e, W = rnd(B, Li, E, seed=9), rnd(1, d, E, scale=0.2, seed=10)
    bias = torch.randn(d, device='cuda')
    out = torch.zeros((off + Li + 1) * B, d, dtype=bf16, device='cuda')
    ops.mixed_gemm(e, W, [(off * B, Li, B, 0, 0)], out, flags=OT_EPI_BIAS, bias=bias, a_transposed_events=True)
    ref = (e.float() @ W[0].float().t() + bias).transpose(0, 1).reshape(Li * B, d)
    close(out[off * B:(off + Li) * B], ref)
    assert out[:off * B].abs().max() == 0 and out[(off + Li) * B:].abs().max() == 0


def test_gemm_rejects_unsupported_shapes():
    A, W = rnd(128, 72), rnd(1, 64, 72)
    with pytest.raises(_lib.OneTransLibraryError, match='K=72'):
        ops.mixed_gemm(A, W, [(0, 1, 128, 0, 0)], torch.empty(128, 64, dtype=bf16, device='cuda'))
    A, W = rnd(128, 64), rnd(1, 40, 64)
    with pytest.raises(_lib.OneTransLibraryError, match='N=40'):
        ops.mixed_gemm(A, W, [(0, 1, 128, 0, 0)], torch.empty(128, 40, dtype=bf16, device='cuda'))


@pytest.mark.parametrize('B,n_s,n_ns', [(32, 9, 4), (256, 20, 3), (2048, 6, 2)])
def test_wgrad_grouped(B, n_s, n_ns):
    M, N = 256, 512
    rows = (n_s + n_ns) * B
    P, Q = rnd(rows, M, seed=11), rnd(rows, N, seed=12)
    C = torch.zeros(1 + n_ns, M, N, device='cuda')
    segs = [(0, 1, n_s * B, 0, 0), (n_s * B, n_ns, B, 1, 1)]
    ops.wgrad_rows(P, Q, segs, C, M * N, N, 1)
    ref = torch.zeros_like(C)
    ref[0] = P[:n_s * B].float().t() @ Q[:n_s * B].float()
    for j in range(n_ns):
        sl = slice((n_s + j) * B, (n_s + j + 1) * B)
        ref[1 + j] = P[sl].float().t() @ Q[sl].float()
    err = ((C - ref).abs().max() / ref.abs().max()).item()
    assert err < 2e-3, err
    # accumulates (does not overwrite)
    ops.wgrad_rows(P, Q, segs, C, M * N, N, 1)
    assert ((C - 2 * ref).abs().max() / ref.abs().max()).item() < 2e-3


@pytest.mark.parametrize('B,n_s,n_ns,M,N,bn', [(32, 9, 4, 256, 512, 0), (200, 7, 3, 128, 1024, 0), (2048, 3, 2, 1024, 256, 0),
                                                (96, 5, 2, 256, 128, 128), (40, 3, 1, 128, 64, 64)])
def test_wgrad_with_bias_gradient(B, n_s, n_ns, M, N, bn):
    """The same pass also delivers the Dense bias gradient: column sums of Q per weight group."""
    rows = (n_s + n_ns) * B
    P, Q = rnd(rows, M, seed=21), rnd(rows, N, seed=22)
    C = torch.zeros(1 + n_ns, M, N, device='cuda')
    db = torch.zeros(1 + n_ns, N, device='cuda')
    segs = [(0, 1, n_s * B, 0, 0), (n_s * B, n_ns, B, 1, 1)]
    wsegs = [dict(P=P[r0:], p_stride_row=M, p_stride_unit=rpu * M, Q=Q[r0:], q_stride_row=N, q_stride_unit=rpu * N,
                  n_units=nu, rows_per_unit=rpu, group_start=g0, group_stride=gs) for (r0, nu, rpu, g0, gs) in segs]
    ops.wgrad(wsegs, C, M, N, M * N, N, 1, block_n=bn, q_colsum=db, q_colsum_group_stride=N)
    ref, dref = torch.zeros_like(C), torch.zeros_like(db)
    ref[0] = P[:n_s * B].float().t() @ Q[:n_s * B].float()
    dref[0] = Q[:n_s * B].float().sum(0)
    for j in range(n_ns):
        sl = slice((n_s + j) * B, (n_s + j + 1) * B)
        ref[1 + j] = P[sl].float().t() @ Q[sl].float()
        dref[1 + j] = Q[sl].float().sum(0)
    assert ((C - ref).abs().max() / ref.abs().max()).item() < 2e-3
    assert ((db - dref).abs().max() / dref.abs().max()).item() < 1e-4


def attn_ref(q, k, v, B, H, Lq, Lk, dh):
    d = H * dh
    q4 = q.float().view(Lq, B, H, dh).permute(1, 2, 0, 3)
    k4 = k.float().view(Lk, B, H, dh).permute(1, 2, 0, 3)
    v4 = v.float().view(Lk, B, H, dh).permute(1, 2, 0, 3)
    s = q4 @ k4.transpose(-1, -2) / math.sqrt(dh)
    qi = torch.arange(Lq, device='cuda')[:, None] + (Lk - Lq)
    ki = torch.arange(Lk, device='cuda')[None, :]
    s = torch.where(ki <= qi, s, torch.full_like(s, -1e9))       # OT/model.py:109-110
    p = torch.softmax(s, -1)
    o = (p @ v4).permute(2, 0, 1, 3).reshape(Lq * B, d)
    lse = torch.logsumexp(s, -1)                                  # [B, H, Lq]
    return o, lse


@pytest.mark.parametrize('B,H,dh,Lq,Lk', [(3, 4, 64, 202, 288), (2, 4, 64, 458, 544), (5, 4, 64, 13, 27), (1, 4, 64, 8, 13),
                                           (2, 4, 96, 224, 288), (2, 2, 64, 1, 1), (1, 4, 64, 1024, 2048), (3, 4, 64, 373, 458),
                                           (40, 4, 64, 300, 300), (2, 4, 64, 640, 700),
                                           # many one-step tiles per CTA (110 per CTA, two slots): the shapes that dead-locked forward v4's
                                           # (sample, head) ring against its shared epilogue
                                           (4096, 4, 64, 8, 24), (4096, 4, 64, 24, 40)])
def test_attention_forward_backward(B, H, dh, Lq, Lk):
    d = H * dh
    q = rnd(Lq * B, d, seed=13)
    kv = rnd(Lk * B, 2 * d, seed=14)
    do = rnd(Lq * B, d, seed=15)
    o = torch.empty(Lq * B, d, dtype=bf16, device='cuda')
    lse = torch.empty(B * H * Lq, device='cuda')
    ops.attn_fwd(q, kv[:, :d], kv[:, d:], o, lse, B, H, Lq, Lk, dh)
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, kv[:, :d], kv[:, d:]))
    o_ref, lse_ref = attn_ref(qf, kf, vf, B, H, Lq, Lk, dh)
    close(o, o_ref)
    assert (lse.view(B, H, Lq) - lse_ref).abs().max().item() < 2e-3
    (o_ref * do.float()).sum().backward()
    dq = torch.empty_like(q)
    dkv = torch.empty_like(kv)
    delta = torch.empty(B * H * Lq, device='cuda')
    ops.attn_bwd(q, kv[:, :d], kv[:, d:], o, lse, do, dq, dkv[:, :d], dkv[:, d:], delta, B, H, Lq, Lk, dh)
    for got, ref in ((dq, qf.grad), (dkv[:, :d], kf.grad), (dkv[:, d:], vf.grad)):
        # relative L2 error; the floor covers gradients that are identically zero (a single key: dS == 0)
        assert ((got.float() - ref).norm() / (ref.norm() + 1e-3 * do.float().norm())).item() < 2e-2


@pytest.mark.parametrize('B,H,Lq,Lk', [(3, 8, 129, 129), (2, 4, 256, 256), (2, 4, 257, 300), (7, 1, 130, 515), (1, 2, 512, 513), (4, 4, 96, 1000),
                                       (2, 4, 193, 193), (3, 4, 385, 400)])
def test_attention_tile_boundaries(B, H, Lq, Lk):
    """Shapes that sit on the edges of the head_dim-64 kernels' tiling: a second query tile of one row (forward v3 pair with an almost
    empty tile B), exact multiples of 128, no causal offset (Lq == Lk), one head, eight heads, a short query tail over a long key
    range (backward items of a single step), an odd number of query tiles.  Same checks as test_attention_forward_backward."""
    test_attention_forward_backward(B, H, 64, Lq, Lk)


def _random_attention_shape(seed):
    """Seeded shape for test_attention_random_shapes: query tails from 1 to 700 rows over key ranges up to 1200, batch sizes chosen
    so that a CTA sees from one to a few hundred (sample, head) items, bounded by the fp32 reference's score tensor (64 M elements)."""
    import random
    r = random.Random(1000 + seed)
    H = r.choice([1, 2, 4, 4, 4, 8])
    Lq = r.choice([r.randint(1, 40), r.randint(41, 128), r.randint(129, 256), r.randint(257, 700)])
    Lk = Lq + r.choice([0, r.randint(0, 16), r.randint(17, 200), r.randint(0, 500)])
    cap = max(1, (64 << 20) // (H * Lq * Lk))
    B = min(cap, r.choice([1, r.randint(2, 40), r.randint(100, 600), r.randint(1000, 5000)]))
    return B, H, Lq, Lk


@pytest.mark.parametrize('seed', range(32))
def test_attention_random_shapes(seed):
    """Shape sweep over the head_dim-64 attention kernels (forward v3 / v4 / v5 by tile count, backward v2): every role of these
    kernels waits on hand-rolled mbarrier rings, and a protocol error can hide behind the shapes one happens to test (forward v4
    dead-locked only with > 100 one-step tiles per CTA).  Same checks as test_attention_forward_backward."""
    B, H, Lq, Lk = _random_attention_shape(seed)
    test_attention_forward_backward(B, H, 64, Lq, Lk)


@pytest.mark.parametrize('B,H,Lq,Lk', [(3, 4, 202, 288), (2, 4, 17, 21), (1, 8, 300, 300)])
def test_attention_head_dim_32(B, H, Lq, Lk):
    """The reference's example scripts use hidden_dim 128 with 4 heads (OT/model.py:420-442, OT/examples/train_example.py:22-27):
    head_dim 32, one 64-byte-swizzle slab per tile.  Same checks as test_attention_forward_backward."""
    test_attention_forward_backward(B, H, 32, Lq, Lk)


@pytest.mark.parametrize('B,Lq,Lk', [(296, 458, 544), (160, 373, 458), (512, 117, 202)])
def test_attention_repeatable(B, Lq, Lk):
    """Race detector for the hand-rolled mbarrier pipelines of the attention kernels (forward v3, backward v2): many (sample, head)
    items per CTA, repeated launches.  The forward and dK / dV have a fixed summation order, so every repeat must be BIT-identical
    to the first; dQ is accumulated with bf16 reductions in item order and may only differ at the rounding level."""
    H, dh = 4, 64
    d = H * dh
    q = rnd(Lq * B, d, seed=31)
    kv = rnd(Lk * B, 2 * d, seed=32)
    do = rnd(Lq * B, d, seed=33)
    outs = []
    for rep in range(6):
        o = torch.empty(Lq * B, d, dtype=bf16, device='cuda')
        lse = torch.empty(B * H * Lq, device='cuda')
        ops.attn_fwd(q, kv[:, :d], kv[:, d:], o, lse, B, H, Lq, Lk, dh)
        dq = torch.empty_like(q)
        dkv = torch.full_like(kv, float('nan'))
        delta = torch.empty(B * H * Lq, device='cuda')
        ops.attn_bwd(q, kv[:, :d], kv[:, d:], o, lse, do, dq, dkv[:, :d], dkv[:, d:], delta, B, H, Lq, Lk, dh)
        outs.append((o, lse, dkv, dq))
    o0, lse0, dkv0, dq0 = outs[0]
    assert torch.isfinite(dkv0.float()).all() and torch.isfinite(dq0.float()).all()
    for o, lse, dkv, dq in outs[1:]:
        assert torch.equal(o, o0) and torch.equal(lse, lse0)
        assert torch.equal(dkv, dkv0)
        assert ((dq.float() - dq0.float()).norm() / dq0.float().norm()).item() < 1e-2


def test_unsupported_head_dim_is_an_error():
    q = rnd(8, 64, seed=1)
    o = torch.empty_like(q)
    lse = torch.empty(2 * 4 * 4, device='cuda')
    with pytest.raises(_lib.OneTransLibraryError, match='head_dim=16'):
        ops.attn_fwd(q, q, q, o, lse, 2, 4, 4, 4, 16)


def test_attention_full_size_properties():
    """BASELINE config 2, layer 0 (B=2048, H=4, Lq=458, Lk=544): with V == 1 every output is exactly 1 (softmax rows
    sum to one) and lse matches a sampled fp32 reference."""
    B, H, dh, Lq, Lk = 2048, 4, 64, 458, 544
    d = H * dh
    q = rnd(Lq * B, d, seed=16)
    kv = rnd(Lk * B, 2 * d, seed=17)
    kv[:, d:] = 1.0
    o = torch.empty(Lq * B, d, dtype=bf16, device='cuda')
    lse = torch.empty(B * H * Lq, device='cuda')
    ops.attn_fwd(q, kv[:, :d], kv[:, d:], o, lse, B, H, Lq, Lk, dh)
    assert (o.float() - 1.0).abs().max().item() < 1e-2
    bsel = torch.tensor([0, 777, 2047], device='cuda')
    qs = q.view(Lq, B, d)[:, bsel].reshape(Lq * 3, d)
    ks = kv[:, :d].reshape(Lk, B, d)[:, bsel].reshape(Lk * 3, d)
    _, lse_ref = attn_ref(qs, ks, ks, 3, H, Lq, Lk, dh)
    assert (lse.view(B, H, Lq)[bsel] - lse_ref).abs().max().item() < 2e-3


@pytest.mark.parametrize('rows,d', [(1000, 256), (333, 384), (64, 1024)])
def test_rmsnorm_forward_backward(rows, d):
    x, dy, dres = rnd(rows, d, seed=18), rnd(rows, d, seed=19), rnd(rows, d, seed=20)
    g = torch.rand(d, device='cuda') + 0.5
    y = torch.empty_like(x)
    rstd = torch.empty(rows, device='cuda')
    ops.rmsnorm_fwd(x, g, y, rstd, 1e-6)
    xf = x.float().requires_grad_(True)
    gf = g.clone().requires_grad_(True)
    yr = xf * torch.rsqrt(xf.square().mean(-1, keepdim=True) + 1e-6) * gf     # OT/model.py:19-23
    close(y, yr)
    (yr * dy.float()).sum().backward()
    dx = torch.empty_like(x)
    dg = torch.zeros(d, device='cuda')
    ops.rmsnorm_bwd(dy, x, rstd, g, dx, dg, dres=dres)
    close(dx, xf.grad + dres.float())
    assert ((dg - gf.grad).abs().max() / gf.grad.abs().max()).item() < 1e-2


def test_ns_tokenizer_fill_rows_colsum():
    B, L_ns, d, nf, row0 = 100, 5, 256, 11, 300
    x = torch.randn(B, nf, device='cuda') * 30      # raw id-like magnitudes
    W = torch.randn(nf, L_ns * d, device='cuda') * 0.05
    b = torch.randn(L_ns * d, device='cuda')
    out = torch.zeros(row0 + L_ns * B, d, dtype=bf16, device='cuda')
    ops.ns_tokenizer_fwd(x, W, b, out, row0, B, L_ns, d)
    ref = (x @ W + b).view(B, L_ns, d).transpose(0, 1).reshape(L_ns * B, d)     # OT/model.py:211-214 -> token-major
    close(out[row0:], ref, 1e-2)
    dout = rnd(row0 + L_ns * B, d, seed=21)
    dW, db = torch.zeros_like(W), torch.zeros_like(b)
    ops.ns_tokenizer_bwd(x, dout, dW, db, row0, B, L_ns, d)
    g = dout[row0:].float().view(L_ns, B, d).transpose(0, 1).reshape(B, L_ns * d)
    assert ((dW - x.t() @ g).abs().max() / (x.t() @ g).abs().max()).item() < 1e-4
    assert ((db - g.sum(0)).abs().max() / g.sum(0).abs().max()).item() < 1e-4
    vec = torch.randn(d, device='cuda')
    ops.fill_rows(vec, out, 10, 7)
    assert torch.equal(out[10:17], vec.to(bf16).expand(7, d)) and out[:10].abs().max() == 0 and out[17:row0].abs().max() == 0
    inp = rnd(900, 384, seed=22)
    o2 = torch.zeros(4, 384, device='cuda')
    ops.colsum(inp, [(0, 1, 500, 0, 0), (500, 2, 200, 2, 1)], o2, 384)
    ref = torch.stack([inp[:500].float().sum(0), torch.zeros(384, device='cuda'), inp[500:700].float().sum(0), inp[700:].float().sum(0)])
    assert ((o2 - ref).abs().max() / ref.abs().max()).item() < 1e-4


def test_full_size_gemm_c2_ffn():
    """BASELINE config 2, layer 0 FFN-1 at full size: 458*2048 rows, K=256 -> N=1024, 33 weight groups,
    checked against fp32 matmul on sampled row blocks (bit-for-bit layout check at scale)."""
    B, n_s, n_ns, K, N = 2048, 426, 32, 256, 1024
    rows = (n_s + n_ns) * B
    A, W = rnd(rows, K, seed=23), rnd(33, N, K, scale=0.1, seed=24)
    bias = torch.randn(33, N, device='cuda')
    out = torch.empty(rows, N, dtype=bf16, device='cuda')
    segs = [(0, 1, n_s * B, 0, 0), (n_s * B, n_ns, B, 1, 1)]
    ops.mixed_gemm(A, W, segs, out, flags=OT_EPI_BIAS, bias=bias)
    for r0, g in [(0, 0), (n_s * B - 300, 0), (n_s * B, 1), ((n_s + 31) * B + 1000, 32), (rows - 128, 32), (500000, 0)]:
        sl = slice(r0, r0 + 128)
        close(out[sl], A[sl].float() @ W[g].float().t() + bias[g])


def _ffn_ref(zn, W1, b1, W2, b2, segs, res=None, mask=None, res_hp=None, hp_row0=0):
    """fp32 reference of the fused FFN with the kernel's storage points: h rounded to bf16 before the second product."""
    rows = zn.shape[0]
    pre = torch.zeros(rows, W1.shape[1], device='cuda')
    y = torch.zeros(rows, W2.shape[1], device='cuda')
    for (r0, n_units, rpu, g0, gs) in segs:
        for u in range(n_units):
            sl, g = slice(r0 + u * rpu, r0 + (u + 1) * rpu), g0 + u * gs
            pre[sl] = zn[sl].float() @ W1[g].float().t() + b1[g]
            h = torch.nn.functional.gelu(pre[sl]).to(bf16).float()
            y[sl] = h @ W2[g].float().t() + b2[g]
    if mask is not None:
        y = y * mask
    if res is not None:
        r = res.float().clone()
        if res_hp is not None:
            r[hp_row0:] = res_hp
        y = y + r
    return pre, y


@pytest.mark.parametrize('B,n_s,n_ns,F,opts', [(256, 3, 2, 1024, 'res,norm,pre'), (128, 1, 0, 256, ''), (40, 5, 3, 512, 'res,norm,pre,hp'),
                                                 (200, 2, 2, 1024, 'res,drop,norm,pre,hp'), (384, 4, 0, 1024, 'res,pre'), (64, 0, 4, 256, 'norm')])
def test_ffn_fused_forward(B, n_s, n_ns, F, opts):
    """ot_ffn_fwd (MixedFFN.call + residual + dropout + next RMSNorm in one kernel, OT/model.py:149-163,196-198) against an fp32
    reference and against the two-GEMM path it replaces: shared run + per-token NS runs, full and partial tiles."""
    d, G = 256, 1 + 5
    rows = (n_s + n_ns) * B
    zn = rnd(rows, d, seed=41)
    W1, W2 = rnd(G, F, d, scale=0.06, seed=42), rnd(G, d, F, scale=0.03, seed=43)
    b1, b2 = 0.1 * torch.randn(G, F, device='cuda'), 0.1 * torch.randn(G, d, device='cuda')
    segs = ([(0, 1, n_s * B, 0, 0)] if n_s else []) + ([(n_s * B, n_ns, B, 2, 1)] if n_ns else [])
    res = rnd(rows, d, seed=44) if 'res' in opts else None
    hp = 'hp' in opts and n_ns > 0
    hp0 = n_s * B
    res_hp = (res[hp0:].float() + 1e-3 * torch.randn(rows - hp0, d, device='cuda')) if hp else None
    out_hp = torch.empty_like(res_hp) if hp else None
    pre = torch.full((rows, F), float('nan'), dtype=bf16, device='cuda') if 'pre' in opts else None
    y = torch.full((rows, d), float('nan'), dtype=bf16, device='cuda')
    gain = 1.0 + 0.1 * torch.randn(d, device='cuda')
    nout = torch.full((rows, d), float('nan'), dtype=bf16, device='cuda') if 'norm' in opts else None
    rstd = torch.empty(rows, device='cuda') if 'norm' in opts else None
    drop = (1234, 0.1) if 'drop' in opts else None
    mask = None
    if drop:
        mask = ops.dropout_mask(torch.ones(rows, d, dtype=bf16, device='cuda'), drop[0], drop[1]).float()
    hsave = torch.full((rows, F), float('nan'), dtype=bf16, device='cuda') if 'pre' in opts else None     # training keeps h too
    ops.ffn_fused(zn, W1, b1, W2, b2, segs, y, pre=pre, h=hsave, res=res, res_hp=res_hp, out_hp=out_hp, hp_row0=hp0 if hp else 0,
                  dropout=drop, norm=(nout, gain, rstd, 1e-6) if nout is not None else None)
    pre_ref, y_ref = _ffn_ref(zn, W1, b1, W2, b2, segs, res, mask, res_hp, hp0)
    if pre is not None:
        close(pre, pre_ref)
        close(hsave, torch.nn.functional.gelu(pre_ref))                  # the h tile, stored as the kernel holds it
    assert not torch.isnan(y.float()).any()
    assert ((y.float() - y_ref).abs().max() / y_ref.abs().max()).item() < 1e-2
    if hp:
        assert ((out_hp - y_ref[hp0:]).abs().max() / y_ref.abs().max()).item() < 3e-3
    if nout is not None:
        r_ref = torch.rsqrt(y_ref.square().mean(-1) + 1e-6)
        assert ((rstd - r_ref).abs().max() / r_ref.abs().max()).item() < 2e-3
        n_ref = y_ref * r_ref[:, None] * gain
        assert ((nout.float() - n_ref).abs().max() / n_ref.abs().max()).item() < 1.5e-2
    # the two-GEMM path on the same inputs (same storage points except that it normalises the bf16-rounded y)
    h2 = torch.empty(rows, F, dtype=bf16, device='cuda')
    y2 = torch.empty(rows, d, dtype=bf16, device='cuda')
    ops.mixed_gemm(zn, W1, segs, h2, flags=OT_EPI_BIAS | OT_EPI_GELU, bias=b1)
    from recommend_b200.engine import split_segments
    kw = dict(res_hp=res_hp, out_hp=torch.empty_like(res_hp), hp_row0=hp0) if hp else {}
    ops.mixed_gemm(h2, W2, split_segments(segs, hp0) if hp else segs, y2, flags=OT_EPI_BIAS | (OT_EPI_RESIDUAL if res is not None else 0), bias=b2, res=res,
                   dropout=drop, **kw)
    assert ((y.float() - y2.float()).abs().max() / y_ref.abs().max()).item() < 1e-2


def test_ffn_fused_full_size_and_repeatable():
    """BASELINE config 2 layer-5 size (32 NS tokens x 2048 rows, one weight group each) and a layer-0-like shared run: the kernel is
    deterministic (same bits twice) and equals the two-GEMM path at full size."""
    d, F, B, n_ns = 256, 1024, 2048, 32
    rows = (8 + n_ns) * B
    zn = rnd(rows, d, seed=51)
    W1, W2 = rnd(1 + n_ns, F, d, scale=0.06, seed=52), rnd(1 + n_ns, d, F, scale=0.03, seed=53)
    b1, b2 = 0.1 * torch.randn(1 + n_ns, F, device='cuda'), 0.1 * torch.randn(1 + n_ns, d, device='cuda')
    res = rnd(rows, d, seed=54)
    segs = [(0, 1, 8 * B, 0, 0), (8 * B, n_ns, B, 1, 1)]
    outs = []
    for _ in range(2):
        y = torch.empty(rows, d, dtype=bf16, device='cuda')
        pre = torch.empty(rows, F, dtype=bf16, device='cuda')
        ops.ffn_fused(zn, W1, b1, W2, b2, segs, y, pre=pre, res=res)
        outs.append((y, pre))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    h2 = torch.empty(rows, F, dtype=bf16, device='cuda')
    pre2 = torch.empty(rows, F, dtype=bf16, device='cuda')
    y2 = torch.empty(rows, d, dtype=bf16, device='cuda')
    ops.mixed_gemm(zn, W1, segs, h2, flags=OT_EPI_BIAS | OT_EPI_GELU, bias=b1, out2=pre2)
    ops.mixed_gemm(h2, W2, segs, y2, flags=OT_EPI_BIAS | OT_EPI_RESIDUAL, bias=b2, res=res)
    close(outs[0][1], pre2, 1e-2)                                           # same products, at most a bf16 rounding apart
    assert ((outs[0][0].float() - y2.float()).abs().max() / y2.float().abs().max()).item() < 1e-2


@pytest.mark.parametrize('B,n_s,n_ns', [(256, 3, 2), (40, 5, 3), (100, 0, 4)])
def test_wgrad_with_gelu_of_p(B, n_s, n_ns):
    """dW2 = gelu(pre)^T dy with the GELU applied to the P tiles inside the weight-gradient kernel (p_gelu) equals the plain
    weight gradient of a stored h = gelu(pre)."""
    F, d, G = 512, 256, 1 + 5
    rows = (n_s + n_ns) * B
    pre, dy = rnd(rows, F, seed=61), rnd(rows, d, seed=62)
    segs = ([(0, 1, n_s * B, 0, 0)] if n_s else []) + ([(n_s * B, n_ns, B, 2, 1)] if n_ns else [])
    h = torch.nn.functional.gelu(pre.float()).to(bf16)
    C1 = torch.zeros(G, F, d, device='cuda')
    C2 = torch.zeros(G, F, d, device='cuda')
    cs1, cs2 = torch.zeros(G, d, device='cuda'), torch.zeros(G, d, device='cuda')
    ops.wgrad_rows(h, dy, segs, C1, F * d, d, 1, q_colsum=cs1, q_colsum_group_stride=d)
    ops.wgrad_rows(pre, dy, segs, C2, F * d, d, 1, q_colsum=cs2, q_colsum_group_stride=d, p_gelu=True)
    ref = torch.zeros(G, F, d, device='cuda')
    for (r0, n_units, rpu, g0, gs) in segs:
        for u in range(n_units):
            sl = slice(r0 + u * rpu, r0 + (u + 1) * rpu)
            ref[g0 + u * gs] += h[sl].float().t() @ dy[sl].float()
    assert ((C1 - ref).abs().max() / ref.abs().max()).item() < 2e-3
    assert ((C2 - ref).abs().max() / ref.abs().max()).item() < 1e-2           # tanh-form GELU + bf16 rounding of h inside the kernel
    assert torch.allclose(cs1, cs2, rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize('B,n_s,n_ns,F', [(256, 3, 2, 1024), (40, 5, 3, 512), (128, 1, 0, 256), (200, 0, 4, 1024)])
def test_ffn_fused_backward(B, n_s, n_ns, F):
    """ot_ffn_bwd: dpre = (dy W2^T) o gelu'(pre) and dzn = dpre W1^T in one kernel against an fp32 reference and the two-GEMM path."""
    d, G = 256, 1 + 5
    rows = (n_s + n_ns) * B
    dy, pre = rnd(rows, d, seed=71), rnd(rows, F, seed=72)
    W2b, W1b = rnd(G, F, d, scale=0.05, seed=73), rnd(G, d, F, scale=0.05, seed=74)     # master layouts: W2 [F, d], W1 [d, F]
    segs = ([(0, 1, n_s * B, 0, 0)] if n_s else []) + ([(n_s * B, n_ns, B, 2, 1)] if n_ns else [])
    dpre = torch.full((rows, F), float('nan'), dtype=bf16, device='cuda')
    dzn = torch.full((rows, d), float('nan'), dtype=bf16, device='cuda')
    ops.ffn_fused_bwd(dy, W2b, W1b, pre, segs, dpre, dzn)
    x = pre.float().requires_grad_(True)
    torch.nn.functional.gelu(x).sum().backward()
    dpre_ref = ref_rows_gemm(dy, W2b, segs) * x.grad
    close(dpre, dpre_ref)
    dzn_ref = ref_rows_gemm(dpre_ref.to(bf16), W1b, segs)
    assert not torch.isnan(dzn.float()).any()
    assert ((dzn.float() - dzn_ref).abs().max() / dzn_ref.abs().max()).item() < 1e-2
    d2, z2 = torch.empty_like(dpre), torch.empty_like(dzn)
    ops.mixed_gemm(dy, W2b, segs, d2, flags=OT_EPI_GELU_GRAD, aux=pre)
    ops.mixed_gemm(d2, W1b, segs, z2)
    close(dpre, d2.float(), 1e-2)
    assert ((dzn.float() - z2.float()).abs().max() / dzn_ref.abs().max()).item() < 1e-2
