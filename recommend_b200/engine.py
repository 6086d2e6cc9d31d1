"""Forward / backward of the OneTrans block and tokenizer on token-major 2-D activations.

Everything here enqueues kernels of ``libonetrans_sm100.so`` (through ``ops``); torch only allocates the
buffers.  Shapes: ``x`` is ``[cur*B, d]`` bf16 with ``row = position*B + sample``; the block keeps the last
``keep`` positions (pyramid tail, OT/model.py:351-371) and returns ``[keep*B, d]``.

Backward accumulates parameter gradients straight into the fp32 ``.grad`` buffers of the parameters
(weight-gradient kernels use fp32 atomics), so a flat gradient buffer can be all-reduced as is."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import ops
from ._lib import OT_EPI_BIAS, OT_EPI_GELU, OT_EPI_GELU_GRAD, OT_EPI_RESIDUAL

bf16 = torch.bfloat16
_SAVE_H = __import__('os').environ.get('OT_FFN_SAVE_H', '0') != '0'      # A/B switch: also store h in the fused forward (see ffn_forward)
_FUSED_BWD = __import__('os').environ.get('OT_FFN_FUSED_BWD', '1') != '0'  # A/B switch: one-kernel input-gradient chain


def _grad_buf(p: torch.Tensor) -> torch.Tensor:
    """fp32 gradient buffer of a parameter (created zeroed on first use, then accumulated into)."""
    if p.grad is None:
        p.grad = torch.zeros_like(p, dtype=torch.float32)
    return p.grad


class BlockWeights:
    """bf16 compute copies of one block's parameters, rebuilt when the fp32 masters change.

    Masters use the Keras ``[in, out]`` kernel layout, packed over weight groups (0 = shared, 1+j = NS
    token j):  Wqkv [G, d, 3d] (q | k | v), Wo [d, d], W1 [G, d, F], W2 [G, F, d].
    The forward GEMM wants ``[G, N=out, K=in]`` (transposed copy); the input-gradient GEMM wants
    ``[G, N=in, K=out]``, which is the master layout itself (plain cast)."""

    def __init__(self):
        self.key = None

    def refresh(self, Wqkv, Wo, W1, W2):
        key = (Wqkv._version, Wo._version, W1._version, W2._version, Wqkv.data_ptr(), W1.data_ptr())
        if key == self.key:
            return
        d = Wqkv.shape[1]
        with torch.no_grad():
            self.Wqkv_b = Wqkv.detach().to(bf16)                                   # [G, d, 3d]
            self.Wq_f = self.Wqkv_b[:, :, :d].transpose(1, 2).contiguous()          # [G, d(out), d(in)]
            self.Wkv_f = self.Wqkv_b[:, :, d:].transpose(1, 2).contiguous()         # [G, 2d, d]
            self.Wo_b = Wo.detach().to(bf16).unsqueeze(0)                           # [1, d, d]
            self.Wo_f = self.Wo_b.transpose(1, 2).contiguous()
            self.W1_b = W1.detach().to(bf16)                                       # [G, d, F]
            self.W1_f = self.W1_b.transpose(1, 2).contiguous()                      # [G, F, d]
            self.W2_b = W2.detach().to(bf16)                                       # [G, F, d]
            self.W2_f = self.W2_b.transpose(1, 2).contiguous()                      # [G, d, F]
        self.key = key


def split_segments(segs: Sequence[ops.Seg], row: int) -> List[ops.Seg]:
    """Split a segment list at output row ``row`` so that the rows from ``row`` on start a segment of their own
    (the GEMM epilogue switches to the fp32 residual stream on a segment boundary, which is a tile boundary)."""
    out: List[ops.Seg] = []
    for (r0, n_units, rpu, g0, gs) in segs:
        end = r0 + n_units * rpu
        if row <= r0 or row >= end:
            out.append((r0, n_units, rpu, g0, gs))
        elif n_units == 1:
            out.append((r0, 1, row - r0, g0, gs))
            out.append((row, 1, end - row, g0, gs))
        else:
            assert (row - r0) % rpu == 0, 'split point must fall on a unit boundary'
            u = (row - r0) // rpu
            out.append((r0, u, rpu, g0, gs))
            out.append((row, n_units - u, rpu, g0 + u * gs, gs))
    return out


# set by train.train_step while a data-parallel backward runs: called with the block whose FFN gradients / whose remaining
# gradients are complete (enqueued on the current stream), so that their all-reduce can start under the rest of the backward
after_ffn_backward = None
after_block_backward = None


def _fused_norm_args(norm: Optional[dict], rows: int, d: int, dev):
    """Allocate the outputs of an RMSNorm fused into a GEMM epilogue (north_star item 4) and return the ``norm=`` tuple
    of ops.mixed_gemm, or None when no norm was asked for / the row does not fit one tile."""
    if norm is None or not ops.can_fuse_norm(d):
        return None
    norm['out'] = torch.empty(rows, d, dtype=bf16, device=dev)
    norm['rstd'] = torch.empty(rows, dtype=torch.float32, device=dev)
    return (norm['out'], norm['gain'], norm['rstd'], norm['eps'])


def mha_forward(xn: torch.Tensor, res: Optional[torch.Tensor], w: BlockWeights, B: int, cur: int, keep: int, H: int,
                L_ns: int, alignment: str, kv_prefix: Optional[torch.Tensor] = None, res_hp: Optional[torch.Tensor] = None,
                drop: Optional[Tuple[int, float]] = None, norm: Optional[dict] = None):
    """MixedMHA.call (OT/model.py:76-122) on normalised ``xn [cur*B, d]``; queries for the last ``keep``
    positions.  ``res`` (``[keep*B, d]``) is added to the Wo output (the block's residual, OT/model.py:193).
    ``kv_prefix [Lc*B, 2d]``: cached K|V rows placed in front of the new ones (OT/model.py:95-98).
    Returns (z, saved, z_hp) with z = res + attn_out @ Wo.  ``norm={'gain': g, 'eps': e}``: the Wo epilogue also writes
    ``RMSNorm(z) * g`` (the block's norm2, OT/model.py:196) and its row statistics into ``norm['out']`` / ``norm['rstd']``
    when the shape allows (``ops.can_fuse_norm``); otherwise the dict comes back without ``'out'``."""
    d = xn.shape[1]
    dh = d // H
    dev = xn.device
    rows, rows_t, off = cur * B, keep * B, (cur - keep) * B
    segs_all = ops.position_segments(0, cur, cur, L_ns, alignment, B)
    segs_tail = ops.position_segments(cur - keep, cur, cur, L_ns, alignment, B)
    Lc = 0 if kv_prefix is None else kv_prefix.shape[0] // B
    kv = torch.empty((Lc + cur) * B, 2 * d, dtype=bf16, device=dev)
    if Lc:
        kv[:Lc * B].copy_(kv_prefix)
    ops.mixed_gemm(xn, w.Wkv_f, segs_all, kv[Lc * B:])
    q = torch.empty(rows_t, d, dtype=bf16, device=dev)
    ops.mixed_gemm(xn[off:], w.Wq_f, segs_tail, q)
    o = torch.empty(rows_t, d, dtype=bf16, device=dev)
    lse = torch.empty(B * H * keep, dtype=torch.float32, device=dev)
    ops.attn_fwd(q, kv[:, :d], kv[:, d:], o, lse, B, H, keep, Lc + cur, dh)
    z = torch.empty(rows_t, d, dtype=bf16, device=dev)
    z_hp = None
    nrm = _fused_norm_args(norm, rows_t, d, dev)
    if res_hp is not None and res is not None:
        # Wo is shared, but the NS-token rows (the last rows of the tail) keep an fp32 residual stream: split the row
        # range at that boundary so that they form their own (tile-aligned) segment
        n_hp = min(res_hp.shape[0], rows_t)
        hp0 = rows_t - n_hp
        z_hp = torch.empty(n_hp, d, dtype=torch.float32, device=dev)
        segs_o = split_segments([(0, 1, rows_t, 0, 0)], hp0)
        ops.mixed_gemm(o, w.Wo_f, segs_o, z, flags=OT_EPI_RESIDUAL, res=res, res_hp=res_hp[res_hp.shape[0] - n_hp:], out_hp=z_hp, hp_row0=hp0,
                       dropout=drop, norm=nrm)
    else:
        ops.mixed_gemm(o, w.Wo_f, [(0, 1, rows_t, 0, 0)], z, flags=OT_EPI_RESIDUAL if res is not None else 0, res=res, dropout=drop,
                       norm=nrm)
    return z, (q, kv, o, lse, segs_all, segs_tail), z_hp


def mha_backward(dz: torch.Tensor, xn: torch.Tensor, saved, w: BlockWeights, Wqkv_grad: torch.Tensor, Wo_grad: torch.Tensor,
                 B: int, cur: int, keep: int, H: int) -> torch.Tensor:
    """Gradients of mha_forward w.r.t. xn (returned, ``[cur*B, d]``) and the weights (accumulated).
    The residual path is the caller's business.  (No kv_prefix support: training never uses a cache.)"""
    q, kv, o, lse, segs_all, segs_tail = saved
    d = xn.shape[1]
    dh = d // H
    dev = xn.device
    rows, rows_t, off = cur * B, keep * B, (cur - keep) * B
    # Wo:  attn_out = o @ Wo
    ops.wgrad_rows(o, dz, [(0, 1, rows_t, 0, 0)], Wo_grad, 0, d, 1)
    do = torch.empty(rows_t, d, dtype=bf16, device=dev)
    ops.mixed_gemm(dz, w.Wo_b, [(0, 1, rows_t, 0, 0)], do)
    # attention: dq | dk | dv land in ONE [rows, 3d] buffer (dq in the tail rows, zero in the pruned head rows), so that the
    # projections' weight gradient and input gradient are one launch each over K = 3d / N = 3d: xn and the gradient rows are
    # read once, dxn is written once (it used to be written by the K|V product and read / re-written by the Q product)
    dqkv = torch.empty(rows, 3 * d, dtype=bf16, device=dev)
    if off > 0:
        dqkv[:off, :d].zero_()
    delta = torch.empty(B * H * keep, dtype=torch.float32, device=dev)
    ops.attn_bwd(q, kv[:, :d], kv[:, d:], o, lse, do, dqkv[off:, :d], dqkv[:, d:2 * d], dqkv[:, 2 * d:], delta, B, H, keep, cur, dh)
    # projections:  q = xn_tail @ Wq[g],  k|v = xn @ Wkv[g]
    ops.wgrad_rows(xn, dqkv, segs_all, Wqkv_grad, d * 3 * d, 3 * d, 1)
    dxn = torch.empty(rows, d, dtype=bf16, device=dev)
    ops.mixed_gemm(dqkv, w.Wqkv_b, segs_all, dxn)
    return dxn


def ffn_forward(zn: torch.Tensor, res: Optional[torch.Tensor], w: BlockWeights, b1: torch.Tensor, b2: torch.Tensor,
                segs: Sequence[ops.Seg], save: bool, res_hp: Optional[torch.Tensor] = None,
                drop: Optional[Tuple[int, float]] = None, norm: Optional[dict] = None):
    """MixedFFN.call (OT/model.py:149-163): ``gelu(zn W1 + b1) W2 + b2`` (+ res).  ``norm`` as in mha_forward: the FFN-2
    epilogue also writes the NEXT block's ``norm1(y)`` (OT/model.py:191)."""
    rows_t, d = zn.shape
    F = w.W1_f.shape[1]
    dev = zn.device
    y = torch.empty(rows_t, d, dtype=bf16, device=dev)
    y_hp = None
    hp = res_hp is not None and res is not None
    if hp:
        hp0 = rows_t - res_hp.shape[0]          # the fp32 stream covers the last rows (the NS tokens)
        y_hp = torch.empty(res_hp.shape[0], d, dtype=torch.float32, device=dev)
    nrm = _fused_norm_args(norm, rows_t, d, dev)
    if ops.can_fuse_ffn(d, F):
        # one kernel: h = gelu(zn W1 + b1) stays on chip, only the pre-activation (the backward's GELU' input) is stored
        pre = torch.empty(rows_t, F, dtype=bf16, device=dev) if save else None
        # OT_FFN_SAVE_H=1 also stores h (straight from the kernel's h tile by TMA) so that dW2 reads it instead of rebuilding
        # gelu(pre) inside the weight-gradient kernel; measured a wash on B200 (forward + 0.8 ms, dW2 - 0.8 ms) for 6 GB more
        # saved activations, so it is off by default
        h = torch.empty(rows_t, F, dtype=bf16, device=dev) if (save and _SAVE_H) else None
        ops.ffn_fused(zn, w.W1_f, b1, w.W2_f, b2, split_segments(segs, hp0) if hp else segs, y, pre=pre, h=h, res=res,
                      res_hp=res_hp if hp else None, out_hp=y_hp, hp_row0=hp0 if hp else 0, dropout=drop, norm=nrm)
        return y, (pre, h), y_hp
    h = torch.empty(rows_t, F, dtype=bf16, device=dev)
    pre = torch.empty(rows_t, F, dtype=bf16, device=dev) if save else None
    ops.mixed_gemm(zn, w.W1_f, segs, h, flags=OT_EPI_BIAS | OT_EPI_GELU, bias=b1, out2=pre)
    flags = OT_EPI_BIAS | (OT_EPI_RESIDUAL if res is not None else 0)
    if hp:
        ops.mixed_gemm(h, w.W2_f, split_segments(segs, hp0), y, flags=flags, bias=b2, res=res, res_hp=res_hp, out_hp=y_hp, hp_row0=hp0,
                       dropout=drop, norm=nrm)
    else:
        ops.mixed_gemm(h, w.W2_f, segs, y, flags=flags, bias=b2, res=res, dropout=drop, norm=nrm)
    return y, (pre, h), y_hp


def ffn_backward(dy: torch.Tensor, zn: torch.Tensor, saved, w: BlockWeights, segs: Sequence[ops.Seg], W1_grad, b1_grad,
                 W2_grad, b2_grad) -> torch.Tensor:
    """Gradient w.r.t. zn (returned) and the FFN parameters (accumulated)."""
    pre, h = saved
    rows_t, d = zn.shape
    F = pre.shape[1]
    dev = zn.device
    # dW2 and db2.  After the fused forward (h kept on chip) the weight-gradient kernel rebuilds h = gelu(pre) tile by tile
    ops.wgrad_rows(h if h is not None else pre, dy, segs, W2_grad, F * d, d, 1, q_colsum=b2_grad, q_colsum_group_stride=d,
                   p_gelu=h is None)
    dpre = torch.empty(rows_t, F, dtype=bf16, device=dev)
    dzn = torch.empty(rows_t, d, dtype=bf16, device=dev)
    if ops.can_fuse_ffn(d, F) and _FUSED_BWD:
        # one kernel: dpre = (dy W2^T) o gelu'(pre) stays on chip as the operand of dzn = dpre W1^T and is stored once for dW1
        ops.ffn_fused_bwd(dy, w.W2_b, w.W1_b, pre, segs, dpre, dzn)
    else:
        ops.mixed_gemm(dy, w.W2_b, segs, dpre, flags=OT_EPI_GELU_GRAD, aux=pre)
        ops.mixed_gemm(dpre, w.W1_b, segs, dzn)
    ops.wgrad_rows(zn, dpre, segs, W1_grad, d * F, F, 1, q_colsum=b1_grad, q_colsum_group_stride=F)  # dW1 and db1
    return dzn


def block_forward(x: torch.Tensor, P: Dict[str, torch.Tensor], w: BlockWeights, B: int, cur: int, keep: int, H: int,
                  L_ns: int, alignment: str, eps: float, save: bool, kv_prefix: Optional[torch.Tensor] = None,
                  x_hp: Optional[torch.Tensor] = None, drop: Optional[Tuple[int, int, float]] = None,
                  pre_norm: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, next_gain: Optional[torch.Tensor] = None):
    """OneTransBlock.call (OT/model.py:186-200) + tail keep (:371).  P: norm1, norm2, b1, b2 (fp32).
    RMSNorms ride in the epilogue of the GEMM that produces their input (north_star item 4): norm2 in the Wo GEMM,
    and - when ``next_gain`` (the next block's norm1 scale) is given - the next block's norm1 in the FFN-2 GEMM; the
    result comes back as 5th value ``(xn_next, rstd_next)`` and enters the next call as ``pre_norm``.
    ``x_hp``: optional fp32 copy of the NS-token rows of ``x`` (its last ``x_hp.shape[0]`` rows) — the
    high-precision residual stream of DESIGN.md §5; returns the matching ``y_hp`` as 4th value.
    ``drop = (seed_attention, seed_ffn, rate)``: Keras inverted dropout on the two branch outputs (OT/model.py:193,198),
    fused into the Wo / FFN-2 epilogues; the masks are recomputed from the seeds in the backward pass."""
    rows, d = x.shape
    assert rows == cur * B
    dev = x.device
    rows_t, off = keep * B, (cur - keep) * B
    if x_hp is not None and x_hp.shape[0] == 0:
        x_hp = None
    if pre_norm is not None:
        xn, r1 = pre_norm                                             # written by the previous block's FFN-2 epilogue
    else:
        xn = torch.empty(rows, d, dtype=bf16, device=dev)
        r1 = torch.empty(rows, dtype=torch.float32, device=dev)
        ops.rmsnorm_fwd(x, P['norm1'], xn, r1, eps, x_hp, rows - (x_hp.shape[0] if x_hp is not None else 0))   # OT/model.py:191
    d_att = (drop[0], drop[2]) if drop is not None else None
    d_ffn = (drop[1], drop[2]) if drop is not None else None
    n2 = {'gain': P['norm2'], 'eps': eps}
    z, mha_saved, z_hp = mha_forward(xn, x[off:], w, B, cur, keep, H, L_ns, alignment, kv_prefix, x_hp, d_att, n2)   # :192-193, :196
    if 'out' in n2:
        zn, r2 = n2['out'], n2['rstd']
    else:
        zn = torch.empty(rows_t, d, dtype=bf16, device=dev)
        r2 = torch.empty(rows_t, dtype=torch.float32, device=dev)
        ops.rmsnorm_fwd(z, P['norm2'], zn, r2, eps, z_hp, rows_t - (z_hp.shape[0] if z_hp is not None else 0))   # :196
    segs_tail = mha_saved[5]
    n1 = {'gain': next_gain, 'eps': eps} if next_gain is not None else None
    y, ffn_saved, y_hp = ffn_forward(zn, z, w, P['b1'], P['b2'], segs_tail, save, z_hp, d_ffn, n1)     # :197-198
    nxt = (n1['out'], n1['rstd']) if n1 is not None and 'out' in n1 else None
    kv = mha_saved[1]
    saved = (x, xn, r1, mha_saved, z, zn, r2, ffn_saved, drop) if save else None
    return y, kv, saved, y_hp, nxt


def block_backward(dy: torch.Tensor, saved, Pm: Dict[str, torch.Tensor], w: BlockWeights, B: int, cur: int, keep: int,
                   H: int, dy_masked: Optional[torch.Tensor] = None, prev_drop: Optional[Tuple[int, float]] = None,
                   on_ffn_done=None):
    """Backward of block_forward.  Pm maps names to the fp32 master parameters (their .grad buffers
    receive the gradients).  Returns ``(dx [cur*B, d], dx_masked)``.
    The dropout masks of the two branches are applied to gradients as they are produced: ``dz`` leaves the norm2
    backward both plain and masked; with ``prev_drop = (seed_ffn, rate)`` of the block below, ``dx`` leaves the norm1
    backward also as that block's masked ``dy`` (returned as 2nd value, passed to its call as ``dy_masked``)."""
    x, xn, r1, mha_saved, z, zn, r2, ffn_saved, drop = saved
    rows, d = x.shape
    dev = x.device
    rows_t, off = keep * B, (cur - keep) * B
    segs_tail = mha_saved[5]
    if not dy.is_contiguous():
        dy = dy.contiguous()
    # y = z + drop(FFN(norm2(z)))
    if drop is None:
        dy_f = dy
    elif dy_masked is not None:
        dy_f = dy_masked
    else:
        dy_f = ops.dropout_mask(dy, drop[1], drop[2])
    dzn = ffn_backward(dy_f, zn, ffn_saved, w, segs_tail, _grad_buf(Pm['W1']), _grad_buf(Pm['b1']), _grad_buf(Pm['W2']),
                       _grad_buf(Pm['b2']))
    if on_ffn_done is not None:
        on_ffn_done()                  # W1, b1, W2, b2 gradients are final: data parallel starts reducing them now
    dz = torch.empty(rows_t, d, dtype=bf16, device=dev)
    # z = x_tail + drop(MHA(norm1(x)))
    dz_a = torch.empty(rows_t, d, dtype=bf16, device=dev) if drop is not None else dz
    ops.rmsnorm_bwd(dzn, z, r2, Pm['norm2'].detach(), dz, _grad_buf(Pm['norm2']), dres=dy,
                    drop_out=(dz_a, drop[0], drop[2], 0) if drop is not None else None)
    dxn = mha_backward(dz_a, xn, mha_saved, w, _grad_buf(Pm['Wqkv']), _grad_buf(Pm['Wo']), B, cur, keep, H)
    dx = torch.empty(rows, d, dtype=bf16, device=dev)
    dx_m = torch.empty(rows, d, dtype=bf16, device=dev) if prev_drop is not None else None
    g1, dg1 = Pm['norm1'].detach(), _grad_buf(Pm['norm1'])
    if off > 0:
        ops.rmsnorm_bwd(dxn[:off], x[:off], r1[:off], g1, dx[:off], dg1, dres=None,
                        drop_out=(dx_m[:off], prev_drop[0], prev_drop[1], 0) if prev_drop is not None else None)
    ops.rmsnorm_bwd(dxn[off:], x[off:], r1[off:], g1, dx[off:], dg1, dres=dz,
                    drop_out=(dx_m[off:], prev_drop[0], prev_drop[1], off) if prev_drop is not None else None)
    return dx, dx_m


# ---------------------------------------------------------------------------------------------------
# tokenizer
# ---------------------------------------------------------------------------------------------------


def tokenizer_forward(ns_x: Optional[torch.Tensor], seq_list: Sequence[Optional[torch.Tensor]], B: int, d: int, L_ns: int,
                      Ws_f: Sequence[torch.Tensor], bs: Sequence[torch.Tensor], sep: torch.Tensor, Wns: torch.Tensor,
                      bns: torch.Tensor):
    """Tokenizer.call (OT/model.py:224-277) written token-major: rows of sequence i, its [SEP] row, ...,
    then the L_NS non-sequence tokens (S first, NS last, OT/model.py:235).  ``seq_list[i]`` is the bf16
    ``[B, L_i, E]`` events of configured sequence i or None when absent."""
    n_seq = len(seq_list)
    layout = []  # (kind, index, first position, length)
    pos = 0
    for i, e in enumerate(seq_list):
        if e is None:
            continue
        layout.append(('seq', i, pos, e.shape[1]))
        pos += e.shape[1]
        if i < n_seq - 1:                                          # OT/model.py:269
            layout.append(('sep', i, pos, 1))
            pos += 1
    L_s = pos
    L = L_s + L_ns
    dev = sep.device
    X0 = torch.empty(L * B, d, dtype=bf16, device=dev)
    for kind, i, p0, n in layout:
        if kind == 'seq':
            ops.mixed_gemm(seq_list[i], Ws_f[i], [(p0 * B, n, B, 0, 0)], X0, flags=OT_EPI_BIAS, bias=bs[i],
                           a_transposed_events=True)                # OT/model.py:265
        else:
            ops.fill_rows(sep.reshape(-1), X0, p0 * B, B)           # OT/model.py:270-272
    X_hp = None
    if ns_x is None:
        X0[L_s * B:].zero_()                                        # OT/model.py:249-251
    elif L_ns > 0:
        X_hp = torch.empty(L_ns * B, d, dtype=torch.float32, device=dev)   # fp32 residual stream of the NS rows
        ops.ns_tokenizer_fwd(ns_x, Wns, bns, X0, L_s * B, B, L_ns, d, X_hp)  # OT/model.py:211-214,253-254
    return X0, L, layout, X_hp


def tokenizer_backward(dX0: torch.Tensor, ns_x, seq_list, layout, B: int, d: int, L_ns: int, Ws, bs, sep, Wns, bns,
                       want_event_grads: Sequence[int] = ()) -> Dict[int, torch.Tensor]:
    """Accumulate the tokenizer's parameter gradients from dX0.  The scalar inputs are data; for the sequences listed
    in ``want_event_grads`` (events that came out of an ``EventEmbedding``) the event gradients
    ``dX0_rows @ Ws[i]^T`` are returned as ``{i: [B, L_i, E]}`` (a transposed view of a token-major buffer)."""
    if not dX0.is_contiguous():
        dX0 = dX0.contiguous()
    L_s = sum(n for _, _, _, n in layout)
    d_events: Dict[int, torch.Tensor] = {}
    for kind, i, p0, n in layout:
        rows = dX0[p0 * B:(p0 + n) * B]
        if kind == 'seq':
            e = seq_list[i]
            E = e.shape[2]
            # dWs[i][k, m] += sum_{l,b} e[b,l,k] * dX0[(p0+l)B+b, m]  -> C[m, k] written transposed
            ops.wgrad([dict(P=rows, p_stride_row=dX0.stride(0), p_stride_unit=B * dX0.stride(0), Q=e, q_stride_row=n * E,
                            q_stride_unit=E, n_units=n, rows_per_unit=B, group_start=0, group_stride=0)],
                      _grad_buf(Ws[i]), d, E, 0, 1, d)
            ops.colsum(dX0, [(p0 * B, 1, n * B, 0, 0)], _grad_buf(bs[i]), 0)
            if i in want_event_grads:
                de = torch.empty(n * B, E, dtype=bf16, device=dX0.device)                  # rows (l, b), token-major
                ops.mixed_gemm(rows, Ws[i].detach().to(bf16).unsqueeze(0), [(0, 1, n * B, 0, 0)], de)   # [E, d] is W[N, K] as stored
                d_events[i] = de.view(n, B, E).transpose(0, 1)
        else:
            ops.colsum(dX0, [(p0 * B, 1, B, 0, 0)], _grad_buf(sep), 0)
    if ns_x is not None:
        ops.ns_tokenizer_bwd(ns_x, dX0, _grad_buf(Wns), _grad_buf(bns), L_s * B, B, L_ns, d)
    return d_events


# ---------------------------------------------------------------------------------------------------
# inference with a cross-candidate cache of the sequence-side K/V  (north_star item 5, PAPER:144-151)
# ---------------------------------------------------------------------------------------------------


def layer_plan(L0: int, L_ns: int, keep_lens: Sequence[int]):
    """Per layer: (cur, Tn, cur_S, keep, Tq, keep_S) — current length, NS tokens alive, S tokens alive, kept
    tail, NS tokens that query, S tokens that query.  The NS tokens are the last rows of the sequence."""
    plan, cur = [], L0
    for keep in keep_lens:
        Tn = min(L_ns, cur)
        Tq = min(Tn, keep)
        plan.append((cur, Tn, cur - Tn, keep, Tq, keep - Tq))
        cur = keep
    return plan


def user_cache_forward(x_s: torch.Tensor, blocks, plan, H: int, eps: float):
    """Stage 1, once per user (B = 1): run the S tokens alone through the stack and keep every layer's K|V.
    Valid because S rows never see NS rows (causal mask, S first: OT/model.py:109-110, 235).
    blocks: list of (P dict with norm1/norm2/b1/b2, BlockWeights).  Returns [kv_l or None]."""
    cache = []
    dev = x_s.device
    d = x_s.shape[1]
    for (P, w), (cur, Tn, cur_S, keep, Tq, keep_S) in zip(blocks, plan):
        if cur_S == 0:
            cache.append(None)
            x_s = x_s[:0]
            continue
        assert x_s.shape[0] == cur_S
        if keep_S > 0:
            x_s, kv, _, _, _ = block_forward(x_s, P, w, 1, cur_S, keep_S, H, 0, 'tail', eps, False)
        else:
            xn = torch.empty(cur_S, d, dtype=bf16, device=dev)
            ops.rmsnorm_fwd(x_s, P['norm1'], xn, None, eps)
            kv = torch.empty(cur_S, 2 * d, dtype=bf16, device=dev)
            ops.mixed_gemm(xn, w.Wkv_f, [(0, 1, cur_S, 0, 0)], kv)
            x_s = x_s[:0]
        cache.append(kv)
    return cache


def extend_user_cache(x_new: torch.Tensor, blocks, plan, cache, H: int, eps: float):
    """Cross-request incremental update of a user's cache (PAPER:151): ``x_new [n, d]`` are the tokens of
    n new behaviours at the tail of the S block.  Streaming rule (restated in oracle ``two_stage_extend``): layer l gains the
    K|V of the new tokens alive at l; a new token attends every key the layer already holds plus the new ones up to itself -
    this is the pyramid kernel's own shape, a query tail over a longer key sequence, with the old keys supplied as
    ``kv_prefix`` instead of being recomputed; ``n_{l+1} = min(n_l, keep_S(l))`` tokens move on.  Returns the new layer list."""
    out = list(cache)
    dev, d = x_new.device, x_new.shape[1]
    for l, ((P, w), (cur, Tn, cur_S, keep, Tq, keep_S)) in enumerate(zip(blocks, plan)):
        n = x_new.shape[0]
        if out[l] is None or n == 0:
            break
        n_q = min(n, keep_S)
        if n_q > 0:
            x_new, out[l], _, _, _ = block_forward(x_new, P, w, 1, n, n_q, H, 0, 'tail', eps, False, kv_prefix=out[l])
        else:
            xn = torch.empty(n, d, dtype=bf16, device=dev)
            ops.rmsnorm_fwd(x_new, P['norm1'], xn, None, eps)
            kv = torch.empty(out[l].shape[0] + n, 2 * d, dtype=bf16, device=dev)
            kv[:out[l].shape[0]].copy_(out[l])
            ops.mixed_gemm(xn, w.Wkv_f, [(0, 1, n, 0, 0)], kv[out[l].shape[0]:])
            out[l] = kv
            x_new = x_new[:0]
    return out


def candidates_forward(x_ns: torch.Tensor, C_: int, blocks, plan, cache, H: int, L_ns: int, eps: float,
                       x_hp: Optional[torch.Tensor] = None):
    """Stage 2: C candidates of the cached user.  x_ns: token-major [L_ns*C, d] NS tokens.  Per layer only the NS
    rows are normalised / projected / fed through the FFN; attention reads the shared S-side K|V from the cache."""
    dev = x_ns.device
    d = x_ns.shape[1]
    dh = d // H
    x = x_ns
    pre_norm = None                      # norm1 of this layer, written by the previous layer's FFN epilogue when it could be fused
    n_layers = len(blocks)
    for l, ((P, w), (cur, Tn, cur_S, keep, Tq, keep_S), kv_s) in enumerate(zip(blocks, plan, cache)):
        assert x.shape[0] == Tn * C_
        rows, rows_t, off = Tn * C_, Tq * C_, (Tn - Tq) * C_
        segs_all = ops.position_segments(cur - Tn, cur, cur, L_ns, 'tail', C_)
        segs_tail = ops.position_segments(cur - Tq, cur, cur, L_ns, 'tail', C_)
        if pre_norm is not None:
            xn = pre_norm
        else:
            xn = torch.empty(rows, d, dtype=bf16, device=dev)
            ops.rmsnorm_fwd(x, P['norm1'], xn, None, eps, x_hp, 0)
        kv = torch.empty(rows, 2 * d, dtype=bf16, device=dev)
        ops.mixed_gemm(xn, w.Wkv_f, segs_all, kv)
        q = torch.empty(rows_t, d, dtype=bf16, device=dev)
        ops.mixed_gemm(xn[off:], w.Wq_f, segs_tail, q)
        o = torch.empty(rows_t, d, dtype=bf16, device=dev)
        Ls = 0 if kv_s is None else kv_s.shape[0]
        ops.attn_ns_cached(q, kv[:, :d], kv[:, d:], None if kv_s is None else kv_s[:, :d], None if kv_s is None else kv_s[:, d:],
                           o, C_, H, Tq, Tn, Ls, dh)
        z = torch.empty(rows_t, d, dtype=bf16, device=dev)
        z_hp = None
        if x_hp is not None:   # every candidate row is an NS row: the whole residual stream is fp32
            z_hp = torch.empty(rows_t, d, dtype=torch.float32, device=dev)
            ops.mixed_gemm(o, w.Wo_f, [(0, 1, rows_t, 0, 0)], z, flags=OT_EPI_RESIDUAL, res=x[off:], res_hp=x_hp[off:], out_hp=z_hp, hp_row0=0)
        else:
            ops.mixed_gemm(o, w.Wo_f, [(0, 1, rows_t, 0, 0)], z, flags=OT_EPI_RESIDUAL, res=x[off:])
        zn = torch.empty(rows_t, d, dtype=bf16, device=dev)
        ops.rmsnorm_fwd(z, P['norm2'], zn, None, eps, z_hp, 0)
        # the fused FFN kernel keeps the finished row in TMEM, so the NEXT layer's norm1 (OT/model.py:191) costs it one more pass
        # over TMEM instead of a kernel of its own; the two-GEMM path (d != 256) keeps the stand-alone norm
        n1 = None
        if l + 1 < n_layers and ops.can_fuse_ffn(d, w.W1_f.shape[1]):
            n1 = {'gain': blocks[l + 1][0]['norm1'], 'eps': eps}
        x, _, x_hp = ffn_forward(zn, z, w, P['b1'], P['b2'], segs_tail, False, z_hp, norm=n1)
        pre_norm = n1['out'] if n1 is not None and 'out' in n1 else None
    return x, x_hp
