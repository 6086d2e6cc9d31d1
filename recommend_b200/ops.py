"""Torch-tensor front end of the C-ABI kernels.  PyTorch supplies device memory and the current stream;
every computation below is a kernel of ``libonetrans_sm100.so``.  All activations are bf16 2-D
``[rows, cols]`` tensors whose rows are token-major (``row = position * B + sample``)."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib as L

Seg = Tuple[int, int, int, int, int]  # (row_start, n_units, rows_per_unit, group_start, group_stride)


class KernelProfiler:
    """Optional per-launch CUDA-event timing on the launching stream (bench.py's roofline numbers).
    Each record: (kernel family, shape tag, start event, end event, algorithmic flops, algorithmic bytes, op-minimal bytes).
    ``bytes`` counts every tensor the launch reads or writes once; ``min_bytes`` leaves out second outputs that exist only because
    of how the work is cut into kernels (a saved pre-activation next to the activation, a fused norm's second output)."""

    def __init__(self, only: Optional[Tuple[str, str]] = None):
        self.records = []
        self.only = only          # (family, shape tag): time just these launches (two event records each are not free)

    def wants(self, name: str, tag: str) -> bool:
        return self.only is None or self.only == (name, tag)

    def summary(self):
        """{(family, tag): dict(launches, ms, flops, bytes)} — call after torch.cuda.synchronize()."""
        out = {}
        for rec in self.records:
            fam, tag, e0, e1, fl, by = rec[:6]
            d = out.setdefault((fam, tag), dict(launches=0, ms=0.0, flops=0.0, bytes=0.0, min_bytes=0.0))
            d['launches'] += 1
            d['ms'] += e0.elapsed_time(e1)
            d['flops'] += fl
            d['bytes'] += by
            d['min_bytes'] += rec[6] if len(rec) > 6 and rec[6] else by
        return out


_PROFILER: Optional[KernelProfiler] = None


def set_profiler(p: Optional[KernelProfiler]) -> None:
    global _PROFILER
    _PROFILER = p


def _run(name: str, fn, p, tag: str = '', flops: float = 0.0, nbytes: float = 0.0, n_launch: int = 1, min_bytes: float = 0.0) -> None:
    prof = _PROFILER
    if prof is not None and not prof.wants(name, tag):
        prof = None
    if prof is not None:
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
    L.check(fn(C.byref(p), _stream()), name)
    if prof is not None:
        e1.record()
        prof.records.append((name, tag, e0, e1, flops, nbytes, min_bytes or nbytes))
    L.count_launch(n_launch)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _check_bf16(t: torch.Tensor, name: str) -> None:
    if t.dtype != torch.bfloat16 or not t.is_cuda:
        raise TypeError(f'{name} must be a CUDA bfloat16 tensor, got {t.dtype} on {t.device}')
    if t.dim() < 1 or t.stride(-1) != 1:
        raise ValueError(f'{name} must be contiguous in its last dimension')


def position_segments(p0: int, p1: int, cur: int, L_ns: int, alignment: str, B: int) -> List[Seg]:
    """Weight-group segments for positions ``[p0, p1)`` of a length-``cur`` sequence, rows relative to
    position ``p0``.  Mirrors ``_get_projection_weights`` (OT/model.py:67-74):
    'head_literal' = as written (position i < L_NS uses dedicated[i]); 'tail' = repair D4 (the NS tokens
    are the last L_NS positions of the ORIGINAL sequence and keep their own weights as the pyramid
    shortens the sequence in front of them).  Group 0 = shared, 1+j = dedicated j."""
    segs: List[Seg] = []
    if alignment == 'tail':
        ns_begin = max(cur - L_ns, 0)           # first position that is an NS token
        first_j = L_ns - (cur - ns_begin)       # its NS index
        a, b = p0, min(p1, ns_begin)
        if b > a:
            segs.append(((a - p0) * B, 1, (b - a) * B, 0, 0))
        a, b = max(p0, ns_begin), p1
        if b > a:
            segs.append(((a - p0) * B, b - a, B, 1 + first_j + (a - ns_begin), 1))
    elif alignment == 'head_literal':
        ded_end = min(L_ns, cur)
        a, b = p0, min(p1, ded_end)
        if b > a:
            segs.append(((a - p0) * B, b - a, B, 1 + a, 1))
        a, b = max(p0, ded_end), p1
        if b > a:
            segs.append(((a - p0) * B, 1, (b - a) * B, 0, 0))
    else:
        raise ValueError(f'unknown ns_param_alignment {alignment!r}')
    return segs


def mixed_gemm(A: torch.Tensor, W: torch.Tensor, segs: Sequence[Seg], out: torch.Tensor, *, flags: int = 0,
               bias: Optional[torch.Tensor] = None, res: Optional[torch.Tensor] = None,
               aux: Optional[torch.Tensor] = None, out2: Optional[torch.Tensor] = None,
               row_scale: Optional[torch.Tensor] = None, a_transposed_events: bool = False,
               block_n: int = 0, swizzle: int = 0, res_hp: Optional[torch.Tensor] = None,
               out_hp: Optional[torch.Tensor] = None, hp_row0: int = 0, dropout: Optional[Tuple[int, float]] = None,
               norm: Optional[Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor], float]] = None) -> torch.Tensor:
    """``out[row] = epilogue(A[row] @ W[group(row)].T)``.  W: ``[G, N, K]`` bf16 (rows may be strided).
    With ``a_transposed_events`` A is ``[B, L_i, K]`` and output rows are ``(l, b)`` token-major.
    ``norm=(norm_out, gain, rstd_out, eps)`` also writes ``RMSNorm(out) * gain`` (and the row statistics the backward
    needs) from the same epilogue; N <= 256 only (``can_fuse_norm``)."""
    _check_bf16(W, 'W')
    _check_bf16(out, 'out')
    G, N, K = W.shape
    p = L.GemmParams()
    p.A = A.data_ptr()
    if a_transposed_events:
        _check_bf16(A, 'A')
        Bsz, Li, Ka = A.shape
        assert Ka == K and A.is_contiguous()
        p.a_dim1, p.a_dim2, p.a_stride1, p.a_stride2, p.a_transposed = Li, Bsz, K, Li * K, 1
    else:
        _check_bf16(A, 'A')
        assert A.dim() == 2 and A.shape[1] == K
        p.a_dim1, p.a_dim2, p.a_stride1, p.a_stride2, p.a_transposed = A.shape[0], 1, A.stride(0), 0, 0
    assert W.stride(2) == 1 and W.stride(0) == N * W.stride(1), 'W groups must be evenly strided rows'
    p.n_groups, p.W, p.ldw, p.N, p.K = G, W.data_ptr(), W.stride(1), N, K
    if dropout is not None and dropout[1] > 0.0:
        flags |= L.OT_EPI_DROPOUT
        p.drop_seed, p.drop_rate = dropout[0] & 0xFFFFFFFF, dropout[1]
    p.n_segs, p.flags = len(segs), flags
    for i, s in enumerate(segs):
        sg = p.segs[i]
        sg.row_start, sg.n_units, sg.rows_per_unit, sg.group_start, sg.group_stride = s[:5]
        sg.a_row_start = s[5] if len(s) > 5 else (0 if a_transposed_events else s[0])
    p.out, p.ldo = out.data_ptr(), out.stride(0)
    if out2 is not None:
        _check_bf16(out2, 'out2')
        p.out2, p.ldo2 = out2.data_ptr(), out2.stride(0)
    if res is not None:
        _check_bf16(res, 'res')
        p.res, p.ldr = res.data_ptr(), res.stride(0)
    if aux is not None:
        _check_bf16(aux, 'aux')
        p.aux, p.ldaux = aux.data_ptr(), aux.stride(0)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_cuda
        p.bias = bias.data_ptr()
        p.bias_group_stride = bias.stride(0) if bias.dim() == 2 else 0
    if row_scale is not None:
        assert row_scale.dtype == torch.float32
        p.row_scale = row_scale.data_ptr()
    p.block_n, p.swizzle = block_n, swizzle
    if res_hp is not None:
        # fp32 residual stream of the NS-token rows (output rows >= hp_row0)
        assert out_hp is not None and res_hp.dtype == torch.float32 and out_hp.dtype == torch.float32
        assert res_hp.stride(-1) == 1 and out_hp.stride(-1) == 1 and res_hp.stride(0) == out_hp.stride(0)
        p.res_hp, p.out_hp, p.ld_hp, p.hp_row0 = res_hp.data_ptr(), out_hp.data_ptr(), res_hp.stride(0), hp_row0
    if norm is not None:
        n_out, n_gain, n_rstd, n_eps = norm
        _check_bf16(n_out, 'norm_out')
        assert n_gain.dtype == torch.float32 and n_gain.is_cuda and n_gain.numel() == N and out2 is None
        p.flags = flags = flags | L.OT_EPI_NORM
        p.norm_out, p.ld_norm, p.norm_gain, p.norm_eps = n_out.data_ptr(), n_out.stride(0), n_gain.data_ptr(), n_eps
        if n_rstd is not None:
            assert n_rstd.dtype == torch.float32 and n_rstd.is_contiguous()
            p.norm_rstd = n_rstd.data_ptr()
    rows = sum(s[1] * s[2] for s in segs)
    groups = sum((s[1] if s[4] else 1) for s in segs)
    n_io = 1 + (out2 is not None) + (res is not None) + (aux is not None) + (norm is not None)
    n_min = 1 + (res is not None) + (aux is not None)       # one output (+ the operands the epilogue semantically needs)
    _run('ot_mixed_gemm', L.load().ot_mixed_gemm, p, f'N{N}_K{K}_f{flags}', 2.0 * rows * N * K,
         rows * K * 2.0 + n_io * rows * N * 2.0 + groups * N * K * 2.0,
         min_bytes=rows * K * 2.0 + n_min * rows * N * 2.0 + groups * N * K * 2.0)
    return out


def can_fuse_ffn(d: int, F: int) -> bool:
    """``ot_ffn_fwd`` is built for d == 256 (its [128 x d] output accumulator and two [128 x 128] FFN-1 accumulators fill the
    512 TMEM columns) and F a multiple of its 128-column chunk.  ``OT_FFN_FUSED=0`` selects the two-GEMM path for A/B runs."""
    import os
    return d == 256 and F % 128 == 0 and os.environ.get('OT_FFN_FUSED', '1') != '0'


def ffn_fused(zn: torch.Tensor, W1_f: torch.Tensor, b1: torch.Tensor, W2_f: torch.Tensor, b2: torch.Tensor, segs: Sequence[Seg],
              out: torch.Tensor, *, pre: Optional[torch.Tensor] = None, h: Optional[torch.Tensor] = None, res: Optional[torch.Tensor] = None,
              res_hp: Optional[torch.Tensor] = None, out_hp: Optional[torch.Tensor] = None, hp_row0: int = 0,
              dropout: Optional[Tuple[int, float]] = None,
              norm: Optional[Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor], float]] = None) -> torch.Tensor:
    """``out = res + drop(gelu(zn @ W1[g].T + b1[g]) @ W2[g].T + b2[g])`` in one kernel (MixedFFN.call, OT/model.py:149-163, with
    the residual / dropout of :198); the hidden activation never reaches HBM.  ``W1_f [G, F, d]``, ``W2_f [G, d, F]`` bf16;
    ``pre [rows, F]`` receives the pre-activation for the backward pass and ``h [rows, F]`` (optional) ``gelu(pre)`` for the dW2
    weight gradient - stored by TMA straight from the kernel's own h tile; ``norm`` as in ``mixed_gemm``."""
    for t, n in ((zn, 'zn'), (W1_f, 'W1'), (W2_f, 'W2'), (out, 'out')):
        _check_bf16(t, n)
    G, F, d = W1_f.shape
    assert W2_f.shape == (G, d, F) and zn.dim() == 2 and zn.shape[1] == d
    assert W1_f.stride(2) == 1 and W1_f.stride(0) == F * W1_f.stride(1) and W2_f.stride(2) == 1 and W2_f.stride(0) == d * W2_f.stride(1)
    assert b1.dtype == torch.float32 and b2.dtype == torch.float32 and b1.is_cuda and b2.is_cuda
    p = L.FfnParams()
    p.zn, p.ldzn, p.W1, p.ldw1, p.W2, p.ldw2 = zn.data_ptr(), zn.stride(0), W1_f.data_ptr(), W1_f.stride(1), W2_f.data_ptr(), W2_f.stride(1)
    p.b1, p.b1_group_stride = b1.data_ptr(), (b1.stride(0) if b1.dim() == 2 else 0)
    p.b2, p.b2_group_stride = b2.data_ptr(), (b2.stride(0) if b2.dim() == 2 else 0)
    p.n_groups, p.d, p.F, p.n_segs = G, d, F, len(segs)
    flags = 0
    for i, s in enumerate(segs):
        sg = p.segs[i]
        sg.row_start, sg.n_units, sg.rows_per_unit, sg.group_start, sg.group_stride = s[:5]
        sg.a_row_start = s[0]
    p.out, p.ldo = out.data_ptr(), out.stride(0)
    if pre is not None:
        _check_bf16(pre, 'pre')
        p.pre, p.ldpre = pre.data_ptr(), pre.stride(0)
    if h is not None:
        _check_bf16(h, 'h')
        p.h, p.ldh = h.data_ptr(), h.stride(0)
    if res is not None:
        _check_bf16(res, 'res')
        flags |= L.OT_EPI_RESIDUAL
        p.res, p.ldr = res.data_ptr(), res.stride(0)
    if dropout is not None and dropout[1] > 0.0:
        flags |= L.OT_EPI_DROPOUT
        p.drop_seed, p.drop_rate = dropout[0] & 0xFFFFFFFF, dropout[1]
    if res_hp is not None:
        assert res is not None and out_hp is not None and res_hp.dtype == torch.float32 and out_hp.dtype == torch.float32
        assert res_hp.stride(-1) == 1 and out_hp.stride(-1) == 1 and res_hp.stride(0) == out_hp.stride(0)
        p.res_hp, p.out_hp, p.ld_hp, p.hp_row0 = res_hp.data_ptr(), out_hp.data_ptr(), res_hp.stride(0), hp_row0
    if norm is not None:
        n_out, n_gain, n_rstd, n_eps = norm
        _check_bf16(n_out, 'norm_out')
        assert n_gain.dtype == torch.float32 and n_gain.is_cuda and n_gain.numel() == d
        flags |= L.OT_EPI_NORM
        p.norm_out, p.ld_norm, p.norm_gain, p.norm_eps = n_out.data_ptr(), n_out.stride(0), n_gain.data_ptr(), n_eps
        if n_rstd is not None:
            assert n_rstd.dtype == torch.float32 and n_rstd.is_contiguous()
            p.norm_rstd = n_rstd.data_ptr()
    p.flags = flags
    rows = sum(s[1] * s[2] for s in segs)
    groups = sum((s[1] if s[4] else 1) for s in segs)
    # algorithmic bytes: zn in, y out (+ residual in, norm out), the saved pre-activation, the weights once per group
    n_io = 2 + (res is not None) + (norm is not None)
    n_f = (pre is not None) + (h is not None)
    _run('ot_ffn_fwd', L.load().ot_ffn_fwd, p, f'd{d}_F{F}_f{flags}{"_pre" if pre is not None else ""}{"_h" if h is not None else ""}', 4.0 * rows * d * F,
         n_io * rows * d * 2.0 + n_f * rows * F * 2.0 + groups * 2.0 * d * F * 2.0,
         min_bytes=(2 + (res is not None)) * rows * d * 2.0 + (rows * F * 2.0 if pre is not None else 0.0) + groups * 2.0 * d * F * 2.0)
    return out


def ffn_fused_bwd(dy: torch.Tensor, W2_b: torch.Tensor, W1_b: torch.Tensor, pre: torch.Tensor, segs: Sequence[Seg], dpre: torch.Tensor,
                  dzn: torch.Tensor) -> torch.Tensor:
    """Input-gradient chain of the FFN in one kernel: ``dpre = (dy @ W2[g].T) * gelu'(pre)`` (written, the dW1 weight gradient reads
    it) and ``dzn = dpre @ W1[g].T`` (returned).  ``W2_b [G, F, d]`` / ``W1_b [G, d, F]`` are plain bf16 casts of the masters."""
    for t, n in ((dy, 'dy'), (W2_b, 'W2_b'), (W1_b, 'W1_b'), (pre, 'pre'), (dpre, 'dpre'), (dzn, 'dzn')):
        _check_bf16(t, n)
    G, F, d = W2_b.shape
    assert W1_b.shape == (G, d, F) and dy.dim() == 2 and dy.shape[1] == d and pre.shape == dpre.shape == (dy.shape[0], F)
    assert W2_b.stride(2) == 1 and W2_b.stride(0) == F * W2_b.stride(1) and W1_b.stride(2) == 1 and W1_b.stride(0) == d * W1_b.stride(1)
    p = L.FfnParams()
    p.zn, p.ldzn, p.W1, p.ldw1, p.W2, p.ldw2 = dy.data_ptr(), dy.stride(0), W2_b.data_ptr(), W2_b.stride(1), W1_b.data_ptr(), W1_b.stride(1)
    p.n_groups, p.d, p.F, p.n_segs, p.flags = G, d, F, len(segs), 0
    for i, s in enumerate(segs):
        sg = p.segs[i]
        sg.row_start, sg.n_units, sg.rows_per_unit, sg.group_start, sg.group_stride = s[:5]
        sg.a_row_start = s[0]
    p.out, p.ldo = dzn.data_ptr(), dzn.stride(0)
    p.pre, p.ldpre, p.h, p.ldh = pre.data_ptr(), pre.stride(0), dpre.data_ptr(), dpre.stride(0)
    rows = sum(s[1] * s[2] for s in segs)
    groups = sum((s[1] if s[4] else 1) for s in segs)
    _run('ot_ffn_bwd', L.load().ot_ffn_bwd, p, f'd{d}_F{F}', 4.0 * rows * d * F,
         2.0 * rows * d * 2.0 + 2.0 * rows * F * 2.0 + groups * 2.0 * d * F * 2.0)
    return dzn


def can_fuse_norm(N: int) -> bool:
    """OT_EPI_NORM needs whole rows inside one 128 x N tile."""
    return N in (64, 128, 256)


WSeg = Tuple[torch.Tensor, int, int, torch.Tensor, int, int, int, int, int, int]


def wgrad(segs: Sequence[dict], Cout: torch.Tensor, Mdim: int, Ndim: int, c_group_stride: int, c_stride_m: int,
          c_stride_n: int, *, block_n: int = 0, swizzle: int = 0, target_ctas: int = 0,
          q_colsum: Optional[torch.Tensor] = None, q_colsum_group_stride: int = 0, p_gelu: bool = False) -> None:
    """``Cout[g][m, n] += sum_rows P[row, m] * Q[row, n]`` (``p_gelu``: ``gelu(P[row, m])``, rebuilt tile by tile); each seg is a dict with keys
    P, p_stride_row, p_stride_unit, Q, q_stride_row, q_stride_unit, n_units, rows_per_unit, group_start,
    group_stride.  ``Cout`` is an fp32 tensor (any view); strides in elements.  ``q_colsum`` (fp32, optional) also
    receives ``sum_rows Q[row, n]`` per group - the bias gradient of the same Dense layer, for free."""
    assert Cout.dtype == torch.float32 and Cout.is_cuda
    p = L.WgradParams()
    p.Mdim, p.Ndim, p.n_segs, p.swizzle = Mdim, Ndim, len(segs), swizzle
    for i, s in enumerate(segs):
        sg = p.segs[i]
        _check_bf16(s['P'], 'P')
        _check_bf16(s['Q'], 'Q')
        sg.P, sg.p_stride_row, sg.p_stride_unit = s['P'].data_ptr(), s['p_stride_row'], s['p_stride_unit']
        sg.Q, sg.q_stride_row, sg.q_stride_unit = s['Q'].data_ptr(), s['q_stride_row'], s['q_stride_unit']
        sg.n_units, sg.rows_per_unit, sg.group_start, sg.group_stride = (
            s['n_units'], s['rows_per_unit'], s['group_start'], s['group_stride'])
    p.C = Cout.data_ptr()
    p.c_group_stride, p.c_stride_m, p.c_stride_n = c_group_stride, c_stride_m, c_stride_n
    p.block_n, p.target_ctas = block_n, target_ctas
    if q_colsum is not None:
        assert q_colsum.dtype == torch.float32 and q_colsum.is_cuda
        p.q_colsum, p.q_colsum_group_stride = q_colsum.data_ptr(), q_colsum_group_stride
    p.p_gelu = 1 if p_gelu else 0
    rows = sum(s['n_units'] * s['rows_per_unit'] for s in segs)
    groups = sum((s['n_units'] if s['group_stride'] else 1) for s in segs)
    _run('ot_wgrad', L.load().ot_wgrad, p, f'M{Mdim}_N{Ndim}' + ('_gelu' if p_gelu else ''), 2.0 * rows * Mdim * Ndim,
         rows * (Mdim + Ndim) * 2.0 + groups * Mdim * Ndim * 4.0)


def wgrad_rows(P: torch.Tensor, Q: torch.Tensor, segs: Sequence[Seg], Cout: torch.Tensor, c_group_stride: int,
               c_stride_m: int, c_stride_n: int, q_colsum: Optional[torch.Tensor] = None, q_colsum_group_stride: int = 0,
               p_gelu: bool = False) -> None:
    """Weight gradient for row-aligned 2-D activations ``P [rows, Mdim]`` and ``Q [rows, Ndim]`` whose
    rows are described by the same position segments the forward GEMM used."""
    wsegs = []
    for (row_start, n_units, rpu, g0, gs) in segs:
        wsegs.append(dict(P=P[row_start:], p_stride_row=P.stride(0), p_stride_unit=rpu * P.stride(0),
                          Q=Q[row_start:], q_stride_row=Q.stride(0), q_stride_unit=rpu * Q.stride(0),
                          n_units=n_units, rows_per_unit=rpu, group_start=g0, group_stride=gs))
    wgrad(wsegs, Cout, P.shape[1], Q.shape[1], c_group_stride, c_stride_m, c_stride_n, q_colsum=q_colsum,
          q_colsum_group_stride=q_colsum_group_stride, p_gelu=p_gelu)


def attn_fwd(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, o: torch.Tensor, lse: torch.Tensor, B: int, H: int,
             Lq: int, Lk: int, head_dim: int, swizzle: int = 0) -> None:
    for t, n in ((q, 'q'), (k, 'k'), (v, 'v'), (o, 'o')):
        _check_bf16(t, n)
    assert lse.dtype == torch.float32 and lse.numel() == B * H * Lq
    p = L.AttnParams()
    p.q, p.ldq, p.k, p.ldk, p.v, p.ldv, p.o, p.ldo = (q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0),
                                                      v.data_ptr(), v.stride(0), o.data_ptr(), o.stride(0))
    p.lse = lse.data_ptr()
    p.B, p.H, p.Lq, p.Lk, p.head_dim, p.swizzle = B, H, Lq, Lk, head_dim, swizzle
    pairs = Lq * Lk - Lq * (Lq - 1) / 2.0
    d = H * head_dim
    _run('ot_attn_fwd', L.load().ot_attn_fwd, p, f'Lq{Lq}_Lk{Lk}', 4.0 * B * H * head_dim * pairs,
         (2.0 * Lq + 2.0 * Lk) * B * d * 2.0)


def attn_bwd(q, k, v, o, lse, d_o, dq, dk, dv, delta, B: int, H: int, Lq: int, Lk: int, head_dim: int,
             swizzle: int = 0) -> None:
    for t, n in ((q, 'q'), (k, 'k'), (v, 'v'), (o, 'o'), (d_o, 'd_o'), (dq, 'dq'), (dk, 'dk'), (dv, 'dv')):
        _check_bf16(t, n)
    p = L.AttnParams()
    p.q, p.ldq, p.k, p.ldk, p.v, p.ldv, p.o, p.ldo = (q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0),
                                                      v.data_ptr(), v.stride(0), o.data_ptr(), o.stride(0))
    p.lse, p.delta = lse.data_ptr(), delta.data_ptr()
    p.d_o, p.lddo, p.dq, p.lddq = d_o.data_ptr(), d_o.stride(0), dq.data_ptr(), dq.stride(0)
    p.dk, p.lddk, p.dv, p.lddv = dk.data_ptr(), dk.stride(0), dv.data_ptr(), dv.stride(0)
    p.B, p.H, p.Lq, p.Lk, p.head_dim, p.swizzle = B, H, Lq, Lk, head_dim, swizzle
    pairs = Lq * Lk - Lq * (Lq - 1) / 2.0
    d = H * head_dim
    _run('ot_attn_bwd', L.load().ot_attn_bwd, p, f'Lq{Lq}_Lk{Lk}', 10.0 * B * H * head_dim * pairs,
         (4.0 * Lq + 4.0 * Lk) * B * d * 2.0, n_launch=2)


def attn_ns_cached(q: torch.Tensor, k_own: torch.Tensor, v_own: torch.Tensor, k_shared: Optional[torch.Tensor],
                   v_shared: Optional[torch.Tensor], o: torch.Tensor, C_: int, H: int, Tq: int, Tn: int, Ls: int, head_dim: int) -> None:
    """NS-token queries of C candidates against the user's cached S-side K/V plus their own NS K/V."""
    for t, n in ((q, 'q'), (k_own, 'k_own'), (v_own, 'v_own'), (o, 'o')):
        _check_bf16(t, n)
    p = L.AttnCachedParams()
    p.q, p.ldq, p.k_own, p.ld_own_k, p.v_own, p.ld_own_v = q.data_ptr(), q.stride(0), k_own.data_ptr(), k_own.stride(0), v_own.data_ptr(), v_own.stride(0)
    if Ls > 0:
        _check_bf16(k_shared, 'k_shared')
        _check_bf16(v_shared, 'v_shared')
        p.k_shared, p.ld_shared_k, p.v_shared, p.ld_shared_v = k_shared.data_ptr(), k_shared.stride(0), v_shared.data_ptr(), v_shared.stride(0)
    p.o, p.ldo = o.data_ptr(), o.stride(0)
    p.C, p.H, p.Tq, p.Tn, p.Ls, p.head_dim = C_, H, Tq, Tn, Ls, head_dim
    pairs = Tq * (Ls + Tn) - Tq * (Tq - 1) / 2.0
    d = H * head_dim
    _run('ot_attn_ns_cached_fwd', L.load().ot_attn_ns_cached_fwd, p, f'Tq{Tq}_Ls{Ls}', 4.0 * C_ * H * head_dim * pairs,
         (2.0 * Tq + 2.0 * Tn) * C_ * d * 2.0 + 2.0 * Ls * d * 2.0)


def rmsnorm_fwd(x: torch.Tensor, gain: torch.Tensor, y: torch.Tensor, rstd: Optional[torch.Tensor], eps: float = 1e-6,
                x_hp: Optional[torch.Tensor] = None, hp_row0: int = 0) -> None:
    _check_bf16(x, 'x')
    _check_bf16(y, 'y')
    assert gain.dtype == torch.float32 and gain.is_cuda and gain.is_contiguous()
    p = L.RmsnormParams()
    p.x, p.ldx, p.y, p.ldy, p.gain, p.rstd = x.data_ptr(), x.stride(0), y.data_ptr(), y.stride(0), gain.data_ptr(), _ptr(rstd)
    p.rows, p.d, p.eps = x.shape[0], x.shape[1], eps
    if x_hp is not None and x_hp.shape[0] > 0:
        assert x_hp.dtype == torch.float32 and x_hp.is_contiguous() and x_hp.shape[1] == x.shape[1]
        assert hp_row0 + x_hp.shape[0] == x.shape[0]
        p.x_hp, p.hp_row0 = x_hp.data_ptr(), hp_row0
    _run('ot_rmsnorm_fwd', L.load().ot_rmsnorm_fwd, p, f'd{x.shape[1]}', 3.0 * x.numel(), x.numel() * 4.0 + x.shape[0] * 4.0)


def rmsnorm_bwd(dy: torch.Tensor, x: torch.Tensor, rstd: torch.Tensor, gain: torch.Tensor, dx: torch.Tensor,
                dgain: torch.Tensor, dres: Optional[torch.Tensor] = None,
                drop_out: Optional[Tuple[torch.Tensor, int, float, int]] = None) -> None:
    """``drop_out = (dx_drop, seed, rate, row0)``: also write ``dropout_mask(dx)`` (mask rows ``row0 + r``) in the same pass."""
    for t, n in ((dy, 'dy'), (x, 'x'), (dx, 'dx')):
        _check_bf16(t, n)
    assert dgain.dtype == torch.float32 and rstd.dtype == torch.float32
    p = L.RmsnormParams()
    p.x, p.ldx, p.gain, p.rstd = x.data_ptr(), x.stride(0), gain.data_ptr(), rstd.data_ptr()
    p.dy, p.lddy, p.dx, p.lddx, p.dgain = dy.data_ptr(), dy.stride(0), dx.data_ptr(), dx.stride(0), dgain.data_ptr()
    if dres is not None:
        _check_bf16(dres, 'dres')
        p.dres, p.lddres = dres.data_ptr(), dres.stride(0)
    p.rows, p.d, p.eps = x.shape[0], x.shape[1], 0.0
    if drop_out is not None:
        dd, seed, rate, row0 = drop_out
        _check_bf16(dd, 'dx_drop')
        p.dx_drop, p.lddx_drop, p.drop_row0, p.drop_seed, p.drop_rate = dd.data_ptr(), dd.stride(0), row0, seed & 0xFFFFFFFF, rate
    if p.rows > 0:
        _run('ot_rmsnorm_bwd', L.load().ot_rmsnorm_bwd, p, f'd{x.shape[1]}', 8.0 * x.numel(),
             x.numel() * 2.0 * (3 + (dres is not None) + (drop_out is not None)) + x.shape[0] * 4.0)


def ns_tokenizer_fwd(x: torch.Tensor, W: torch.Tensor, bias: torch.Tensor, out: torch.Tensor, row0: int, B: int,
                     L_ns: int, d: int, out_hp: Optional[torch.Tensor] = None) -> None:
    assert x.dtype == torch.float32 and x.is_contiguous() and W.dtype == torch.float32 and W.is_contiguous()
    p = L.NsTokenizerParams()
    p.x, p.W, p.bias, p.out, p.ldo = x.data_ptr(), W.data_ptr(), bias.data_ptr(), out.data_ptr(), out.stride(0)
    p.row0, p.B, p.L_ns, p.d, p.n_feat = row0, B, L_ns, d, x.shape[1]
    if out_hp is not None:
        assert out_hp.dtype == torch.float32 and out_hp.is_contiguous() and out_hp.shape == (L_ns * B, d)
        p.out_hp = out_hp.data_ptr()
    _run('ot_ns_tokenizer_fwd', L.load().ot_ns_tokenizer_fwd, p, '', 2.0 * B * L_ns * d * x.shape[1], B * L_ns * d * 2.0)


def ns_tokenizer_bwd(x: torch.Tensor, dout: torch.Tensor, dW: torch.Tensor, dbias: torch.Tensor, row0: int, B: int,
                     L_ns: int, d: int) -> None:
    assert x.dtype == torch.float32 and x.is_contiguous() and dW.dtype == torch.float32 and dW.is_contiguous()
    p = L.NsTokenizerParams()
    p.x, p.dout, p.ldo, p.dW, p.dbias = x.data_ptr(), dout.data_ptr(), dout.stride(0), dW.data_ptr(), dbias.data_ptr()
    p.row0, p.B, p.L_ns, p.d, p.n_feat = row0, B, L_ns, d, x.shape[1]
    _run('ot_ns_tokenizer_bwd', L.load().ot_ns_tokenizer_bwd, p, '', 2.0 * B * L_ns * d * x.shape[1], B * L_ns * d * 2.0)


def fill_rows(vec: torch.Tensor, out: torch.Tensor, row0: int, n_rows: int) -> None:
    assert vec.dtype == torch.float32 and vec.is_contiguous()
    L.check(L.load().ot_fill_rows(vec.data_ptr(), out.data_ptr(), out.stride(0), row0, n_rows, out.shape[1], _stream()),
            'ot_fill_rows')
    L.count_launch()


def dropout_mask(inp: torch.Tensor, seed: int, rate: float) -> torch.Tensor:
    """Backward of the epilogue dropout: ``keep ? inp/(1-rate) : 0`` with the mask the forward GEMM used."""
    _check_bf16(inp, 'inp')
    out = torch.empty_like(inp)
    prof = _PROFILER
    if prof is not None and not prof.wants('ot_dropout_mask', f'N{inp.shape[1]}'):
        prof = None
    if prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    L.check(L.load().ot_dropout_mask(inp.data_ptr(), inp.stride(0), out.data_ptr(), out.stride(0), inp.shape[0], inp.shape[1],
                                     seed & 0xFFFFFFFF, rate, _stream()), 'ot_dropout_mask')
    if prof is not None:
        e1.record()
        prof.records.append(('ot_dropout_mask', f'N{inp.shape[1]}', e0, e1, float(inp.numel()), inp.numel() * 4.0))
    L.count_launch()
    return out


def colsum(inp: torch.Tensor, segs: Sequence[Seg], out: torch.Tensor, out_group_stride: int) -> None:
    """``out[group][n] += sum_rows inp[row, n]`` for each unit of each segment (bias gradients)."""
    _check_bf16(inp, 'inp')
    assert out.dtype == torch.float32
    for (row_start, n_units, rpu, g0, gs) in segs:
        p = L.ColsumParams()
        p.in_, p.ld, p.row_start = inp.data_ptr(), inp.stride(0), row_start
        p.n_units, p.rows_per_unit, p.group_start, p.group_stride = n_units, rpu, g0, gs
        p.out, p.out_group_stride, p.N = out.data_ptr(), out_group_stride, inp.shape[1]
        _run('ot_colsum', L.load().ot_colsum, p, f'N{inp.shape[1]}', float(n_units * rpu * inp.shape[1]),
             n_units * rpu * inp.shape[1] * 2.0)


def clip_rmsprop_step(param_ptrs: torch.Tensor, seg_off: torch.Tensor, seg_numel: torch.Tensor, grad: torch.Tensor,
                      rms: torch.Tensor, mom: Optional[torch.Tensor], sqnorm: torch.Tensor, *, lr: float, rho: float,
                      momentum: float, eps: float, clip_norm: float, grad_scale: float = 1.0, zero_grad: bool = False,
                      seg_slot: Optional[torch.Tensor] = None, n_slots: int = 0) -> None:
    """Per-tensor ``clip_by_norm`` then Keras RMSprop on flat fp32 buffers (OT/train.py:131-138).  ``param_ptrs`` /
    ``seg_off`` / ``seg_numel`` are int64 DEVICE tables (see ``ot_rmsprop_params``)."""
    for t in (param_ptrs, seg_off, seg_numel):
        if not t.is_cuda or t.dtype != torch.int64:
            raise RuntimeError('clip_rmsprop_step: tables must be CUDA int64 tensors (no CPU fallback)')
    for t in (grad, rms, sqnorm) + ((mom,) if mom is not None else ()):
        if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
            raise RuntimeError('clip_rmsprop_step: buffers must be contiguous CUDA fp32 tensors (no CPU fallback)')
    p = L.RmspropParams()
    p.param_ptrs, p.seg_off, p.seg_numel = param_ptrs.data_ptr(), seg_off.data_ptr(), seg_numel.data_ptr()
    p.n_seg, p.n_flat = seg_numel.numel(), grad.numel()
    p.grad, p.rms, p.mom, p.sqnorm = grad.data_ptr(), rms.data_ptr(), _ptr(mom), sqnorm.data_ptr()
    p.lr, p.rho, p.momentum, p.eps, p.clip_norm, p.grad_scale = lr, rho, momentum, eps, clip_norm, grad_scale
    p.zero_grad = 1 if zero_grad else 0
    if seg_slot is not None:      # clip slots finer than tensors (one per Keras variable), see ot_rmsprop_params.seg_slot
        if not seg_slot.is_cuda or seg_slot.dtype != torch.int64 or seg_slot.shape != (seg_numel.numel(), 4) or not seg_slot.is_contiguous():
            raise RuntimeError('clip_rmsprop_step: seg_slot must be a contiguous CUDA int64 [n_seg, 4] table')
        if sqnorm.numel() < n_slots:
            raise RuntimeError('clip_rmsprop_step: sqnorm needs one entry per clip slot')
        p.seg_slot, p.n_slots = seg_slot.data_ptr(), n_slots
    n = float(grad.numel())
    _run('ot_clip_rmsprop_step', L.load().ot_clip_rmsprop_step, p, f'n{grad.numel()}', 10.0 * n,
         n * (4.0 * (clip_norm > 0) + 24.0 + 8.0 * (momentum != 0.0) + 4.0 * bool(zero_grad)), n_launch=2 if clip_norm > 0 else 1)


def heads_fwd(x: torch.Tensor, gain: torch.Tensor, eps: float, heads: Sequence[Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]],
              labels: Optional[torch.Tensor] = None):
    """Output norm + task heads (+ BCE) on the last-token rows ``x [B, d]`` (fp32).  ``heads``: per task
    ``(kernel0 [d, h], bias0 [h], kernel1 [h, 1], bias1 [1])``.  Returns ``(probs [T, B], logits [T, B], loss | None, saved)``."""
    if not x.is_cuda or x.dtype != torch.float32 or x.stride(-1) != 1:
        raise RuntimeError('heads_fwd: x must be a CUDA fp32 tensor with contiguous rows (no CPU fallback)')
    B, d = x.shape
    T, Hd = len(heads), heads[0][0].shape[1]
    dev = x.device
    xn = torch.empty(B, d, dtype=torch.float32, device=dev)
    rstd = torch.empty(B, dtype=torch.float32, device=dev)
    pre = torch.empty(T, B, Hd, dtype=torch.float32, device=dev)
    out = torch.empty(2, T, B, dtype=torch.float32, device=dev)          # probs, logits
    p = L.HeadsParams()
    p.x, p.ldx, p.gain, p.eps, p.B, p.d, p.hidden, p.n_tasks = x.data_ptr(), x.stride(0), gain.data_ptr(), eps, B, d, Hd, T
    for t, (k0, b0, k1, b1) in enumerate(heads):
        for w in (k0, b0, k1, b1):
            assert w.dtype == torch.float32 and w.is_cuda and w.is_contiguous()
        p.W0[t], p.b0[t], p.W1[t], p.b1[t] = k0.data_ptr(), b0.data_ptr(), k1.data_ptr(), b1.data_ptr()
    p.xn, p.rstd, p.pre, p.probs, p.logits = xn.data_ptr(), rstd.data_ptr(), pre.data_ptr(), out[0].data_ptr(), out[1].data_ptr()
    loss = g_bce = None
    if labels is not None:
        assert labels.shape == (T, B) and labels.dtype == torch.float32 and labels.is_contiguous()
        lg = torch.zeros(1 + T * B, dtype=torch.float32, device=dev)   # [loss | g_bce]
        loss, g_bce = lg[0], lg[1:].view(T, B)
        p.labels, p.loss, p.g_bce = labels.data_ptr(), loss.data_ptr(), g_bce.data_ptr()
    _run('ot_heads_fwd', L.load().ot_heads_fwd, p, f'd{d}_h{Hd}_T{T}', 2.0 * T * B * d * Hd, 4.0 * B * d * 2 + 4.0 * T * B * Hd)
    return out[0], out[1], loss, (xn, rstd, pre, g_bce)


def heads_bwd(x: torch.Tensor, gain: torch.Tensor, eps: float, heads, saved, dlogit: torch.Tensor, grads, dgain: torch.Tensor) -> torch.Tensor:
    """Backward of ``heads_fwd``: ``dlogit [T, B]`` -> accumulates into ``grads`` (per task ``(dW0, db0, dW1, db1)`` fp32 buffers) and
    ``dgain``; returns ``dx [B, d]`` (fp32)."""
    xn, rstd, pre, _ = saved
    B, d = x.shape
    T, Hd = len(heads), heads[0][0].shape[1]
    dx = torch.empty(B, d, dtype=torch.float32, device=x.device)
    dpre = torch.empty(T, B, Hd, dtype=torch.float32, device=x.device)
    assert dlogit.shape == (T, B) and dlogit.dtype == torch.float32 and dlogit.is_contiguous()
    p = L.HeadsParams()
    p.x, p.ldx, p.gain, p.eps, p.B, p.d, p.hidden, p.n_tasks = x.data_ptr(), x.stride(0), gain.data_ptr(), eps, B, d, Hd, T
    for t, ((k0, b0, k1, b1), (g0, gb0, g1, gb1)) in enumerate(zip(heads, grads)):
        p.W0[t], p.b0[t], p.W1[t], p.b1[t] = k0.data_ptr(), b0.data_ptr(), k1.data_ptr(), b1.data_ptr()
        for gq in (g0, gb0, g1, gb1):
            assert gq.dtype == torch.float32 and gq.is_contiguous()
        p.dW0[t], p.db0[t], p.dW1[t], p.db1[t] = g0.data_ptr(), gb0.data_ptr(), g1.data_ptr(), gb1.data_ptr()
    p.xn, p.rstd, p.pre = xn.data_ptr(), rstd.data_ptr(), pre.data_ptr()
    p.dlogit, p.dpre, p.dgain, p.dx, p.lddx = dlogit.data_ptr(), dpre.data_ptr(), dgain.data_ptr(), dx.data_ptr(), dx.stride(0)
    _run('ot_heads_bwd', L.load().ot_heads_bwd, p, f'd{d}_h{Hd}_T{T}', 6.0 * T * B * d * Hd, 4.0 * B * d * 3 + 8.0 * T * B * Hd, n_launch=2)
    return dx
