"""OneTrans model — the reference's module API (``rank/scaling_up/oneTrans/practice/model.py``,
"OT/model.py") as ``torch.nn.Module``s whose arithmetic runs in hand-written sm_100a kernels.

Same class names, constructor arguments and call signatures as the Keras classes:
``RMSNorm(dim, eps)``, ``MixedMHA(config)``, ``MixedFFN(config)``, ``OneTransBlock(config)``,
``Tokenizer(config)``, ``PyramidScheduler(config)``, ``OneTransModel(config)``,
``create_onetrans_model(model_type)``.  Tensors at module boundaries are ``[B, L, d]`` like the
reference's; physically they are views of token-major ``[L, B, d]`` buffers, so chaining
modules never copies.  Parameters keep the Keras ``[in, out]`` kernel layout, packed over weight groups
(index 0 = shared S-token weights, 1+j = NS token j).

There is no CPU path: every module raises if its input is not on a CUDA device."""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F_

from . import engine, ops
from .config import OneTransConfig, get_model_config
from .schedule import PyramidScheduler, resolve_keep_lens

bf16 = torch.bfloat16


def _glorot_(t: torch.Tensor, fan_in: int, fan_out: int, gen: Optional[torch.Generator] = None) -> torch.Tensor:
    """Keras ``glorot_uniform`` (Dense default)."""
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    with torch.no_grad():
        t.copy_((torch.rand(t.shape, generator=gen, dtype=torch.float32) * 2 - 1) * lim)
    return t


def _require_cuda(t: torch.Tensor, who: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f'{who}: input is on {t.device}; the OneTrans kernels are sm_100a-only and there is no CPU fallback')


def _to_token_major(x: torch.Tensor) -> Tuple[torch.Tensor, int, int]:
    """``[B, L, d]`` (any strides) -> contiguous token-major 2-D ``[L*B, d]`` bf16; free when ``x`` already is
    a ``[B, L, d]`` view of such a buffer."""
    B, L, d = x.shape
    xt = x.transpose(0, 1)
    if xt.dtype != bf16:
        xt = xt.to(bf16)
    if not xt.is_contiguous():
        xt = xt.contiguous()
    return xt.reshape(L * B, d), B, L


def _from_token_major(x2: torch.Tensor, B: int, L: int) -> torch.Tensor:
    return x2.view(L, B, x2.shape[1]).transpose(0, 1)


# ---------------------------------------------------------------------------------------------------
# RMSNorm  (OT/model.py:11-23)
# ---------------------------------------------------------------------------------------------------


class _RMSNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x2, scale, eps):
        y = torch.empty_like(x2)
        rstd = torch.empty(x2.shape[0], dtype=torch.float32, device=x2.device)
        ops.rmsnorm_fwd(x2, scale.detach(), y, rstd, eps)
        ctx.save_for_backward(x2, rstd)
        ctx.scale = scale
        return y

    @staticmethod
    def backward(ctx, dy):
        x2, rstd = ctx.saved_tensors
        dx = torch.empty_like(x2)
        ops.rmsnorm_bwd(dy.contiguous(), x2, rstd, ctx.scale.detach(), dx, engine._grad_buf(ctx.scale))
        return dx, None, None


class RMSNorm(nn.Module):
    """``x * rsqrt(mean(x^2, -1) + eps) * scale`` (OT/model.py:19-23)."""

    def __init__(self, dim: int, eps: float = 1e-6):
        super().__init__()
        self.scale = nn.Parameter(torch.ones(dim))
        self.eps = eps

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        _require_cuda(x, 'RMSNorm')
        shp, dt = x.shape, x.dtype
        x2 = x.reshape(-1, shp[-1]).to(bf16).contiguous()
        y = _RMSNormFn.apply(x2, self.scale, self.eps)
        return y.reshape(shp).to(dt)


# ---------------------------------------------------------------------------------------------------
# MixedMHA / MixedFFN  (OT/model.py:26-163)
# ---------------------------------------------------------------------------------------------------


class MixedMHA(nn.Module):
    """Mixed-parameter causal multi-head attention (OT/model.py:26-122).

    Parameters: ``Wqkv [1+L_NS, d, 3d]`` = (Wq | Wk | Wv) per weight group, no bias (OT/model.py:38-54);
    ``Wo [d, d]`` shared, no bias (:57)."""

    def __init__(self, config: OneTransConfig):
        super().__init__()
        self.config = config
        self.hidden_dim = config.hidden_dim
        self.num_heads = config.num_heads
        self.head_dim = config.hidden_dim // config.num_heads
        self.num_ns_tokens = config.num_ns_tokens
        d, G = self.hidden_dim, 1 + config.num_ns_tokens
        self.Wqkv = nn.Parameter(torch.empty(G, d, 3 * d))
        self.Wo = nn.Parameter(torch.empty(d, d))
        for part in range(3):
            _glorot_(self.Wqkv.data[:, :, part * d:(part + 1) * d], d, d)
        _glorot_(self.Wo.data, d, d)
        # 3*G Keras Dense kernels live in this tensor (q | k | v column blocks per weight group): tf.clip_by_norm treats each
        # one separately (OT/train.py:135); (elements per group, row length, columns per variable) for train.ClipRMSprop
        self.Wqkv._ot_clip_layout = (d * 3 * d, 3 * d, d)

    # reference-style accessors (OT/model.py:38-54)
    def _slice(self, part: int, j: Optional[int]) -> torch.Tensor:
        d = self.hidden_dim
        return self.Wqkv[0 if j is None else 1 + j, :, part * d:(part + 1) * d]

    @property
    def Wq_shared(self): return self._slice(0, None)
    @property
    def Wk_shared(self): return self._slice(1, None)
    @property
    def Wv_shared(self): return self._slice(2, None)

    def forward(self, x: torch.Tensor, training: bool = False,
                kv_cache: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, query_len: Optional[int] = None):
        """x: normalised tokens ``[B, L, d]``.  Returns ``(output [B, Lq, d], (k, v))`` with k, v ``[B, Lk, d]``
        (OT/model.py:76-122).  ``kv_cache=(k, v)``: cached keys/values placed in front (OT/model.py:95-98, with
        the causal mask aligned to the sequence tail, repair D6).  ``query_len``: only the last rows query."""
        _require_cuda(x, 'MixedMHA')
        x2, B, L = _to_token_major(x)
        keep = L if query_len is None else int(query_len)
        prefix = None
        if kv_cache is not None:
            k2, _, Lc = _to_token_major(kv_cache[0])
            v2, _, _ = _to_token_major(kv_cache[1])
            prefix = torch.cat([k2, v2], dim=1)
        out, kv = _MHAFn.apply(x2, self.Wqkv, self.Wo, self, B, L, keep, prefix, torch.is_grad_enabled())
        d = self.hidden_dim
        Lk = kv.shape[0] // B
        k = _from_token_major(kv[:, :d], B, Lk)
        v = _from_token_major(kv[:, d:], B, Lk)
        return _from_token_major(out, B, keep), (k, v)


class _MHAFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x2, Wqkv, Wo, mod: MixedMHA, B, cur, keep, prefix, grad_on=True):
        ctx.set_materialize_grads(False)   # no zero tensors for the (k, v) output nobody differentiates
        w = _weights_of(mod, None)
        cfg = mod.config
        out, saved, _ = engine.mha_forward(x2, None, w, B, cur, keep, cfg.num_heads, cfg.num_ns_tokens,
                                           cfg.ns_param_alignment, prefix)
        kv = saved[1]
        ctx.mark_non_differentiable(kv)
        # ctx.needs_input_grad is True for a Parameter input even under torch.no_grad(), and grad mode is always off inside
        # Function.forward: the caller's grad mode (``grad_on``) decides
        if prefix is not None and grad_on and any(ctx.needs_input_grad):
            raise RuntimeError('MixedMHA: kv_cache is an inference feature; call it under torch.no_grad()')
        ctx.saved = (x2, saved, w, mod, B, cur, keep)
        return out, kv

    @staticmethod
    def backward(ctx, dout, _dkv):
        if dout is None:
            return (None,) * 9
        x2, saved, w, mod, B, cur, keep = ctx.saved
        dxn = engine.mha_backward(dout.contiguous(), x2, saved, w, engine._grad_buf(mod.Wqkv), engine._grad_buf(mod.Wo),
                                  B, cur, keep, mod.config.num_heads)
        return dxn, None, None, None, None, None, None, None, None


class MixedFFN(nn.Module):
    """Mixed-parameter feed-forward (OT/model.py:125-163): ``Dense(F, gelu) -> Dense(d)`` with biases.
    Parameters packed over weight groups: ``W1 [G, d, F]``, ``b1 [G, F]``, ``W2 [G, F, d]``, ``b2 [G, d]``."""

    def __init__(self, config: OneTransConfig):
        super().__init__()
        self.config = config
        self.hidden_dim = config.hidden_dim
        self.ffn_dim = config.ffn_dim
        self.num_ns_tokens = config.num_ns_tokens
        d, Fd, G = config.hidden_dim, config.ffn_dim, 1 + config.num_ns_tokens
        self.W1 = nn.Parameter(_glorot_(torch.empty(G, d, Fd), d, Fd))
        self.b1 = nn.Parameter(torch.zeros(G, Fd))
        self.W2 = nn.Parameter(_glorot_(torch.empty(G, Fd, d), Fd, d))
        self.b2 = nn.Parameter(torch.zeros(G, d))
        for t in (self.W1, self.b1, self.W2, self.b2):      # G Keras variables each (one per weight group), see MixedMHA
            n = t[0].numel()
            t._ot_clip_layout = (n, n, n)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        _require_cuda(x, 'MixedFFN')
        x2, B, L = _to_token_major(x)
        y = _FFNFn.apply(x2, self.W1, self, B, L, torch.is_grad_enabled())
        return _from_token_major(y, B, L)


class _FFNFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x2, W1, mod: MixedFFN, B, cur, grad_on=True):
        w = _weights_of(None, mod)
        cfg = mod.config
        segs = ops.position_segments(0, cur, cur, cfg.num_ns_tokens, cfg.ns_param_alignment, B)
        need_grad = grad_on and any(ctx.needs_input_grad)   # grad mode is always off inside Function.forward
        y, saved, _ = engine.ffn_forward(x2, None, w, mod.b1.detach(), mod.b2.detach(), segs, save=need_grad)
        ctx.saved = (x2, saved, w, mod, segs) if need_grad else None
        return y

    @staticmethod
    def backward(ctx, dy):
        x2, saved, w, mod, segs = ctx.saved
        g = engine._grad_buf
        dx = engine.ffn_backward(dy.contiguous(), x2, saved, w, segs, g(mod.W1), g(mod.b1), g(mod.W2), g(mod.b2))
        return dx, None, None, None, None, None


class _PartialWeights:
    """bf16 compute copies for a stand-alone MixedMHA or MixedFFN (the block keeps a joint cache)."""

    def __init__(self):
        self.key = None


def _weights_of(mha: Optional[MixedMHA], ffn: Optional[MixedFFN]):
    mod = mha if mha is not None else ffn
    w = getattr(mod, '_ot_weights', None)
    if w is None:
        w = _PartialWeights()
        object.__setattr__(mod, '_ot_weights', w)
    if mha is not None:
        key = (mha.Wqkv._version, mha.Wo._version, mha.Wqkv.data_ptr())
        if key != w.key:
            d = mha.hidden_dim
            with torch.no_grad():
                w.Wqkv_b = mha.Wqkv.detach().to(bf16)
                w.Wq_f = w.Wqkv_b[:, :, :d].transpose(1, 2).contiguous()
                w.Wkv_f = w.Wqkv_b[:, :, d:].transpose(1, 2).contiguous()
                w.Wo_b = mha.Wo.detach().to(bf16).unsqueeze(0)
                w.Wo_f = w.Wo_b.transpose(1, 2).contiguous()
            w.key = key
    else:
        key = (ffn.W1._version, ffn.W2._version, ffn.W1.data_ptr())
        if key != w.key:
            with torch.no_grad():
                w.W1_b = ffn.W1.detach().to(bf16)
                w.W1_f = w.W1_b.transpose(1, 2).contiguous()
                w.W2_b = ffn.W2.detach().to(bf16)
                w.W2_f = w.W2_b.transpose(1, 2).contiguous()
            w.key = key
    return w


# ---------------------------------------------------------------------------------------------------
# OneTransBlock  (OT/model.py:166-200)
# ---------------------------------------------------------------------------------------------------


class OneTransBlock(nn.Module):
    """Pre-norm causal block: ``z = x + drop(MHA(norm1(x)))``, ``y = z + drop(FFN(norm2(z)))`` (OT/model.py:186-200)."""

    def __init__(self, config: OneTransConfig):
        super().__init__()
        self.config = config
        self.norm1 = RMSNorm(config.hidden_dim, getattr(config, 'rms_eps', 1e-6))
        self.norm2 = RMSNorm(config.hidden_dim, getattr(config, 'rms_eps', 1e-6))
        self.attention = MixedMHA(config)
        self.ffn = MixedFFN(config)
        self.dropout_rate = config.dropout_rate
        self._w = engine.BlockWeights()

    def _weights(self) -> engine.BlockWeights:
        self._w.refresh(self.attention.Wqkv, self.attention.Wo, self.ffn.W1, self.ffn.W2)
        return self._w

    def _params(self) -> Dict[str, torch.Tensor]:
        return {'norm1': self.norm1.scale, 'norm2': self.norm2.scale, 'Wqkv': self.attention.Wqkv, 'Wo': self.attention.Wo,
                'W1': self.ffn.W1, 'b1': self.ffn.b1, 'W2': self.ffn.W2, 'b2': self.ffn.b2}

    def forward_token_major(self, x2: torch.Tensor, B: int, cur: int, keep: int, training: bool = False,
                            kv_prefix: Optional[torch.Tensor] = None, x_hp: Optional[torch.Tensor] = None,
                            pre_norm=None, next_gain: Optional[torch.Tensor] = None, prev_block=None):
        """Token-major entry used by OneTransModel: ``x2 [cur*B, d]`` -> ``([keep*B, d], kv [Lk*B, 2d], y_hp)``.
        ``x_hp``: fp32 copy of the NS-token rows (high-precision residual stream) or None.
        ``pre_norm`` / ``next_gain``: norm1 of this block already computed by the previous block's FFN-2 epilogue /
        ask this block's FFN-2 epilogue for the next block's norm1 (left in ``self._next_norm``)."""
        drop = None
        if training and self.dropout_rate > 0.0:
            # Keras Dropout(rate) on both branch outputs (OT/model.py:184,193,198); seeds come from torch's CPU generator
            s = torch.randint(0, 2 ** 31 - 1, (2,))
            drop = (int(s[0]), int(s[1]), float(self.dropout_rate))
        object.__setattr__(self, '_last_drop', drop)
        # the block below (if any) gets its masked output gradient from this block's norm1 backward
        prev = None
        if prev_block is not None and getattr(prev_block, '_last_drop', None) is not None:
            prev = (prev_block, (prev_block._last_drop[1], prev_block._last_drop[2]))
        return _BlockFn.apply(x2, self.norm1.scale, self, B, cur, keep, kv_prefix, x_hp, drop, pre_norm, next_gain, prev,
                              torch.is_grad_enabled())

    def forward(self, x: torch.Tensor, training: bool = False,
                kv_cache: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, query_len: Optional[int] = None):
        """``x [B, L, d]`` -> ``(x' [B, Lq, d], (k, v))`` (OT/model.py:186-200).  ``query_len`` keeps only the last
        rows (the pyramid tail, OT/model.py:351-371)."""
        _require_cuda(x, 'OneTransBlock')
        x2, B, L = _to_token_major(x)
        keep = L if query_len is None else int(query_len)
        prefix = None
        if kv_cache is not None:
            k2, _, _ = _to_token_major(kv_cache[0])
            v2, _, _ = _to_token_major(kv_cache[1])
            prefix = torch.cat([k2, v2], dim=1)
        y, kv, _ = self.forward_token_major(x2, B, L, keep, training, prefix)
        d = self.config.hidden_dim
        Lk = kv.shape[0] // B
        return _from_token_major(y, B, keep), (_from_token_major(kv[:, :d], B, Lk), _from_token_major(kv[:, d:], B, Lk))


class _BlockFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x2, anchor, blk: OneTransBlock, B, cur, keep, kv_prefix, x_hp, drop, pre_norm=None, next_gain=None, prev=None,
                grad_on=True):
        # without this autograd fills a [cur*B, 2d] bf16 and a [L_NS*B, d] fp32 zero tensor per block and step for the
        # two outputs nobody differentiates (1.3 ms/step at C2, profiles/README.md)
        ctx.set_materialize_grads(False)
        cfg = blk.config
        # ctx.needs_input_grad is True for the Parameter anchor even under torch.no_grad(), and grad mode is always off inside
        # Function.forward: the caller's grad mode (``grad_on``) decides whether anything is saved (eval / inference forwards
        # keep no activations and skip the GELU pre-activation output)
        need_grad = grad_on and any(ctx.needs_input_grad)
        if need_grad and kv_prefix is not None:
            raise RuntimeError('OneTransBlock: kv_cache is an inference feature; call it under torch.no_grad()')
        w = blk._weights()
        P = {'norm1': blk.norm1.scale.detach(), 'norm2': blk.norm2.scale.detach(), 'b1': blk.ffn.b1.detach(),
             'b2': blk.ffn.b2.detach()}
        y, kv, saved, y_hp, nxt = engine.block_forward(x2, P, w, B, cur, keep, cfg.num_heads, cfg.num_ns_tokens, cfg.ns_param_alignment,
                                                       blk.norm1.eps, need_grad, kv_prefix, x_hp, drop, pre_norm,
                                                       None if next_gain is None else next_gain.detach())
        object.__setattr__(blk, '_next_norm', nxt)    # side channel to the next block (plain tensors, no autograd edge)
        if y_hp is None:
            y_hp = y.new_zeros(0, dtype=torch.float32)
        ctx.mark_non_differentiable(kv, y_hp)
        ctx.saved = (saved, w, blk, B, cur, keep, prev)
        return y, kv, y_hp

    @staticmethod
    def backward(ctx, dy, _dkv, _dhp):
        if dy is None:
            return (None,) * 13
        saved, w, blk, B, cur, keep, prev = ctx.saved
        # masked copy of dy left behind by the block above (its norm1 backward wrote it in the same pass as dy itself)
        dy_masked = None
        m = getattr(blk, '_dy_masked', None)
        if m is not None:
            object.__setattr__(blk, '_dy_masked', None)
            if m[1].data_ptr() == dy.data_ptr() and m[1].shape == dy.shape:
                dy_masked = m[0]
        hook = engine.after_ffn_backward
        dx, dx_m = engine.block_backward(dy, saved, blk._params(), w, B, cur, keep, blk.config.num_heads, dy_masked,
                                         prev[1] if prev is not None else None,
                                         (lambda: hook(blk)) if hook is not None else None)
        if prev is not None:
            object.__setattr__(prev[0], '_dy_masked', (dx_m, dx))
        if engine.after_block_backward is not None:
            engine.after_block_backward(blk)      # data parallel: this block's gradients may start their all-reduce now
        ctx.saved = None
        return dx, None, None, None, None, None, None, None, None, None, None, None, None


# ---------------------------------------------------------------------------------------------------
# Tokenizer  (OT/model.py:203-277)
# ---------------------------------------------------------------------------------------------------


class Tokenizer(nn.Module):
    """Unified tokenizer: non-sequence scalars -> Dense(d*L_NS)+Reshape (OT/model.py:211-214); each behaviour
    sequence ``[B, L_i, 64]`` -> Dense(d) (:217-219), ``[SEP]`` after every present sequence except the last
    configured one (:269-272); output ``concat([S, NS])`` (:235)."""

    def __init__(self, config: OneTransConfig):
        super().__init__()
        self.config = config
        d, L_ns, E = config.hidden_dim, config.num_ns_tokens, getattr(config, 'seq_feature_dim', 64)
        n_feat = len(config.ns_features)
        self.ns_kernel = nn.Parameter(_glorot_(torch.empty(n_feat, d * L_ns), n_feat, d * L_ns))
        self.ns_bias = nn.Parameter(torch.zeros(d * L_ns))
        n_seq = len(config.feature_config['sequence_features'])
        self.seq_kernels = nn.ParameterList([nn.Parameter(_glorot_(torch.empty(E, d), E, d)) for _ in range(n_seq)])
        self.seq_biases = nn.ParameterList([nn.Parameter(torch.zeros(d)) for _ in range(n_seq)])
        self.sep_embedding = nn.Parameter((torch.rand(1, d) - 0.5) * 0.1)     # Keras Embedding default U(-.05,.05)
        self._cache_key = None

    def _seq_weights(self) -> List[torch.Tensor]:
        key = tuple(p._version for p in self.seq_kernels) + (self.seq_kernels[0].data_ptr(),)
        if key != self._cache_key:
            with torch.no_grad():
                self._Ws_f = [p.detach().t().contiguous().to(bf16).unsqueeze(0) for p in self.seq_kernels]  # [1, d, E]
            self._cache_key = key
        return self._Ws_f

    def _gather_inputs(self, non_seq_features: Dict[str, torch.Tensor], seq_features: Dict[str, torch.Tensor]):
        cfg = self.config
        feats = [non_seq_features[n] for n in cfg.ns_features if n in non_seq_features]      # OT/model.py:243-247
        ref = feats[0] if feats else next(iter(seq_features.values()))
        _require_cuda(ref, 'Tokenizer')
        B = ref.shape[0]
        ns_x = None
        if feats:
            if len(feats) != len(cfg.ns_features):
                raise ValueError(f'Tokenizer was built for non-sequence features {cfg.ns_features}; got '
                                 f'{[n for n in cfg.ns_features if n in non_seq_features]} (set config.ns_feature_names)')
            ns_x = torch.cat([f.reshape(B, 1).to(torch.float32) for f in feats], dim=1).contiguous()  # :253 (+D9 cast)
        seq_list = []
        for name in cfg.feature_config['sequence_features']:
            e = seq_features.get(name)
            if e is not None:
                e = e.to(bf16).contiguous()
            seq_list.append(e)
        return ns_x, seq_list, B

    def forward_token_major(self, non_seq_features, seq_features):
        ns_x, seq_list, B = self._gather_inputs(non_seq_features, seq_features)
        X0, L, X_hp = _TokenizerFn.apply(self.ns_kernel, self, ns_x, seq_list, B, *[e for e in seq_list if e is not None])
        self._last_hp = X_hp if X_hp.numel() else None
        return X0, B, L

    def forward(self, non_seq_features: Dict[str, torch.Tensor], seq_features: Dict[str, torch.Tensor]) -> torch.Tensor:
        X0, B, L = self.forward_token_major(non_seq_features, seq_features)
        return _from_token_major(X0, B, L)


class _TokenizerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, tok: Tokenizer, ns_x, seq_list, B, *events):
        # ``events`` repeats the present entries of seq_list as explicit tensor arguments so that autograd can hand a
        # gradient back to an EventEmbedding in front of the tokenizer
        ctx.set_materialize_grads(False)
        cfg = tok.config
        d, L_ns = cfg.hidden_dim, cfg.num_ns_tokens
        X0, L, layout, X_hp = engine.tokenizer_forward(ns_x, seq_list, B, d, L_ns, tok._seq_weights(),
                                                       [b.detach() for b in tok.seq_biases], tok.sep_embedding.detach(),
                                                       tok.ns_kernel.detach(), tok.ns_bias.detach())
        if X_hp is None:
            X_hp = X0.new_zeros(0, dtype=torch.float32)
        ctx.mark_non_differentiable(X_hp)
        ctx.saved = (tok, ns_x, seq_list, layout, B)
        ctx.present = [i for i, e in enumerate(seq_list) if e is not None]
        return X0, L, X_hp

    @staticmethod
    def backward(ctx, dX0, _dL=None, _dhp=None):
        n_ev = len(ctx.present)
        if dX0 is None:
            return (None,) * (5 + n_ev)
        tok, ns_x, seq_list, layout, B = ctx.saved
        cfg = tok.config
        want = [i for j, i in enumerate(ctx.present) if ctx.needs_input_grad[5 + j]]
        d_ev = engine.tokenizer_backward(dX0, ns_x, seq_list, layout, B, cfg.hidden_dim, cfg.num_ns_tokens, list(tok.seq_kernels),
                                         list(tok.seq_biases), tok.sep_embedding, tok.ns_kernel, tok.ns_bias, want)
        return (None, None, None, None, None) + tuple(d_ev.get(i) for i in ctx.present)


# ---------------------------------------------------------------------------------------------------
# OneTransModel  (OT/model.py:305-416)
# ---------------------------------------------------------------------------------------------------


class TaskHead(nn.Module):
    """``Dense(d/2, gelu) -> Dense(1, sigmoid)`` (OT/model.py:327-330), Keras ``[in, out]`` kernels, fp32."""

    def __init__(self, d: int):
        super().__init__()
        self.kernel0 = nn.Parameter(_glorot_(torch.empty(d, d // 2), d, d // 2))
        self.bias0 = nn.Parameter(torch.zeros(d // 2))
        self.kernel1 = nn.Parameter(_glorot_(torch.empty(d // 2, 1), d // 2, 1))
        self.bias1 = nn.Parameter(torch.zeros(1))

    def forward(self, x: torch.Tensor) -> torch.Tensor:  # returns logits
        h = F_.gelu(x @ self.kernel0 + self.bias0)
        return h @ self.kernel1 + self.bias1


class _HeadsFn(torch.autograd.Function):
    """Output norm + task heads (+ BCE) as two library calls (``ot_heads_fwd`` / ``ot_heads_bwd``).  ``x_last`` carries the
    autograd edge into the blocks (bf16 rows of the residual stream); ``x_val`` are the fp32 values actually used (the
    fp32 NS stream when the model keeps one).  Returns (probs [T, B], logits [T, B], loss: scalar, or empty without labels)."""

    @staticmethod
    def forward(ctx, x_last, anchor, model, x_val, labels):
        ctx.set_materialize_grads(False)
        heads = [(h.kernel0.detach(), h.bias0.detach(), h.kernel1.detach(), h.bias1.detach()) for h in model.task_heads.values()]
        probs, logits, loss, saved = ops.heads_fwd(x_val, model.output_norm.scale.detach(), model.output_norm.eps, heads, labels)
        ctx.saved = (model, x_val, heads, saved, probs, x_last.dtype)
        return probs, logits, (loss if loss is not None else probs.new_zeros(0))

    @staticmethod
    def backward(ctx, d_probs, d_logits, d_loss):
        model, x_val, heads, saved, probs, x_dtype = ctx.saved
        dlogit = None
        if d_loss is not None and d_loss.numel() and saved[3] is not None:
            dlogit = saved[3] * d_loss                              # fused BCE: d loss / d logit was written by the forward kernel
        if d_logits is not None:
            dlogit = d_logits if dlogit is None else dlogit + d_logits
        if d_probs is not None:
            dl = d_probs * probs * (1.0 - probs)                    # sigmoid'
            dlogit = dl if dlogit is None else dlogit + dl
        if dlogit is None:
            return None, None, None, None, None
        g = engine._grad_buf
        grads = [(g(h.kernel0), g(h.bias0), g(h.kernel1), g(h.bias1)) for h in model.task_heads.values()]
        dx = ops.heads_bwd(x_val, model.output_norm.scale.detach(), model.output_norm.eps, heads, saved, dlogit.contiguous().float(),
                           grads, g(model.output_norm.scale))
        ctx.saved = None
        return dx.to(x_dtype), None, None, None, None


class OneTransModel(nn.Module):
    """OneTrans unified ranking model (OT/model.py:305-408): tokenizer -> pyramid-stacked blocks -> output
    RMSNorm -> task heads on the last token.  Returns ``{task: probabilities [B, 1]}`` like the reference
    (``return_logits=True`` returns the pre-sigmoid values)."""

    def __init__(self, config: OneTransConfig):
        super().__init__()
        self.config = config
        self.tokenizer = Tokenizer(config)
        self.blocks = nn.ModuleList([OneTransBlock(config) for _ in range(config.num_layers)])
        self.pyramid_scheduler = PyramidScheduler(config)
        self.output_norm = RMSNorm(config.hidden_dim, getattr(config, 'rms_eps', 1e-6))
        self.task_heads = nn.ModuleDict({t: TaskHead(config.hidden_dim) for t in config.tasks})
        self.kv_cache = None

    def forward(self, non_seq_features: Dict[str, torch.Tensor], seq_features: Dict[str, torch.Tensor],
                training: bool = False, use_kv_cache: bool = False, return_logits: bool = False,
                _labels: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        if use_kv_cache and not training:
            # OT/model.py:359-363 (as repaired, D6): reuse the cached sequence-side K/V of the current user; the
            # first call (or a call after reset_kv_cache) builds it from batch-1 sequence features.
            if self.kv_cache is None:
                self.build_kv_cache(seq_features)
            return self.score_candidates(non_seq_features, return_logits=return_logits)
        x2, B, L = self.tokenizer.forward_token_major(non_seq_features, seq_features)       # OT/model.py:342
        keep_lens = resolve_keep_lens(self.config, L)                                      # :349 (+D2, D5)
        cur = L
        x_hp = self.tokenizer._last_hp if getattr(self.config, 'hp_ns_residual', True) else None
        pre_norm = None
        n_blocks = len(self.blocks)
        for i, (block, keep) in enumerate(zip(self.blocks, keep_lens)):                    # :348
            if x_hp is not None:   # the NS rows that survive into this layer's input (suffix of the fp32 stream)
                x_hp = x_hp[x_hp.shape[0] - min(self.config.num_ns_tokens, cur) * B:]
            next_gain = self.blocks[i + 1].norm1.scale if i + 1 < n_blocks else None
            x2, _, x_hp = block.forward_token_major(x2, B, cur, keep, training, None, x_hp, pre_norm, next_gain,
                                                    self.blocks[i - 1] if i > 0 else None)                      # :366-371
            pre_norm = block._next_norm
            object.__setattr__(block, '_next_norm', None)
            if x_hp.numel() == 0:
                x_hp = None
            cur = keep
        x_last = x2[(cur - 1) * B:cur * B]
        return self._heads(x_last, return_logits, x_hp[-B:] if x_hp is not None else None, _labels)

    def forward_with_loss(self, non_seq_features: Dict[str, torch.Tensor], seq_features: Dict[str, torch.Tensor],
                          labels: Dict[str, torch.Tensor], training: bool = True):
        """``(loss, probabilities)`` of one training step's forward: the model call of OT/train.py:118 and the loss of
        :124-128 (sum over tasks of Keras BinaryCrossentropy) in one pass - the BCE is evaluated inside the head kernel."""
        tasks = list(self.task_heads.keys())
        y = torch.stack([labels[t].reshape(-1).to(torch.float32) for t in tasks]).contiguous()
        out = self.forward(non_seq_features, seq_features, training=training, _labels=y)
        return out.pop('_loss'), out

    def _heads(self, x_last: torch.Tensor, return_logits: bool, x_val: Optional[torch.Tensor] = None,
               labels: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """Output norm (OT/model.py:384) and heads (:388-391) on the last token, in fp32 (``ot_heads_fwd``).  ``x_val``: fp32
        values of the rows (the fp32 NS stream) when they differ from ``x_last``, which then only carries the gradient."""
        if x_val is None:
            x_val = x_last.detach().float()
        probs, logits, loss = _HeadsFn.apply(x_last, self.output_norm.scale, self, x_val.contiguous(), labels)
        src = logits if return_logits else probs
        out = {task: src[t].unsqueeze(1) for t, task in enumerate(self.task_heads.keys())}
        if labels is not None:
            out['_loss'] = loss
        return out

    # ---- inference with a cross-candidate cache of the sequence-side K/V (PAPER:144-151; repair D6) ----
    def _block_bundles(self):
        return [({'norm1': b.norm1.scale.detach(), 'norm2': b.norm2.scale.detach(), 'b1': b.ffn.b1.detach(), 'b2': b.ffn.b2.detach()},
                 b._weights()) for b in self.blocks]

    @torch.no_grad()
    def build_kv_cache(self, seq_features: Dict[str, torch.Tensor]):
        """Stage 1, once per user: ``seq_features`` with batch 1.  Runs the S tokens alone through the stack and
        returns (and stores in ``self.kv_cache``) every layer's sequence-side K|V."""
        cfg = self.config
        if cfg.ns_param_alignment != 'tail':
            raise NotImplementedError("the KV cache needs ns_param_alignment='tail' (NS weights on the NS tokens)")
        tok = self.tokenizer
        seq_list = []
        for name in cfg.feature_config['sequence_features']:
            e = seq_features.get(name)
            if e is not None:
                _require_cuda(e, 'build_kv_cache')
                if e.shape[0] != 1:
                    raise ValueError('build_kv_cache caches ONE user: sequence features must have batch 1')
                e = e.to(bf16).contiguous()
            seq_list.append(e)
        x_s, L_s, _, _ = engine.tokenizer_forward(None, seq_list, 1, cfg.hidden_dim, 0, tok._seq_weights(),
                                               [b.detach() for b in tok.seq_biases], tok.sep_embedding.detach(),
                                               tok.ns_kernel.detach(), tok.ns_bias.detach())
        L0 = L_s + cfg.num_ns_tokens
        plan = engine.layer_plan(L0, cfg.num_ns_tokens, resolve_keep_lens(cfg, L0))
        layers = engine.user_cache_forward(x_s, self._block_bundles(), plan, cfg.num_heads, self.blocks[0].norm1.eps)
        self.kv_cache = {'plan': plan, 'layers': layers, 'L0': L0, 'has_last_sequence': seq_list[-1] is not None, 'appended': 0}
        return self.kv_cache

    @torch.no_grad()
    def extend_kv_cache(self, new_events: torch.Tensor, kv_cache=None):
        """Cross-request incremental update (PAPER:151): ``new_events [1, n, 64]`` are n new behaviours of the cached user in the
        LAST configured sequence (the only place where new tokens land at the tail of the S block; a new event in an earlier
        sequence shifts every token behind it and needs ``build_kv_cache``).  Only the n new tokens are computed; each layer's
        key set grows by the new tokens alive there (``engine.extend_user_cache``).  With the pyramid off this equals a fresh
        build on the longer sequence; with it on, which OLD tokens a layer holds stays as decided at build time."""
        cache = kv_cache if kv_cache is not None else self.kv_cache
        if cache is None:
            raise RuntimeError('extend_kv_cache: no KV cache; call build_kv_cache(seq_features) first')
        if not cache.get('has_last_sequence', False):
            raise ValueError('extend_kv_cache: the cache was built without the last sequence; new events would not be at the tail')
        cfg, tok = self.config, self.tokenizer
        _require_cuda(new_events, 'extend_kv_cache')
        if new_events.dim() != 3 or new_events.shape[0] != 1 or new_events.shape[2] != cfg.seq_feature_dim:
            raise ValueError(f'extend_kv_cache: new_events must be [1, n, {cfg.seq_feature_dim}], got {tuple(new_events.shape)}')
        if new_events.shape[1] == 0:
            return cache
        n_seq = len(cfg.feature_config['sequence_features'])
        seq_list = [None] * (n_seq - 1) + [new_events.to(bf16).contiguous()]
        x_new, _, _, _ = engine.tokenizer_forward(None, seq_list, 1, cfg.hidden_dim, 0, tok._seq_weights(),
                                                 [b.detach() for b in tok.seq_biases], tok.sep_embedding.detach(),
                                                 tok.ns_kernel.detach(), tok.ns_bias.detach())
        cache['layers'] = engine.extend_user_cache(x_new, self._block_bundles(), cache['plan'], cache['layers'], cfg.num_heads,
                                                   self.blocks[0].norm1.eps)
        cache['appended'] = cache.get('appended', 0) + new_events.shape[1]
        return cache

    @torch.no_grad()
    def score_candidates(self, non_seq_features: Dict[str, torch.Tensor], kv_cache=None, return_logits: bool = False):
        """Stage 2: score C candidates of the cached user; only the NS tokens are computed per candidate."""
        cache = kv_cache if kv_cache is not None else self.kv_cache
        if cache is None:
            raise RuntimeError('score_candidates: no KV cache; call build_kv_cache(seq_features) first')
        cfg = self.config
        tok = self.tokenizer
        ns_x, _, C_ = tok._gather_inputs(non_seq_features, {})
        if ns_x is None:
            raise ValueError('score_candidates needs the non-sequence features of the candidates')
        d, L_ns = cfg.hidden_dim, cfg.num_ns_tokens
        x_ns = torch.empty(L_ns * C_, d, dtype=bf16, device=ns_x.device)
        x_hp = torch.empty(L_ns * C_, d, dtype=torch.float32, device=ns_x.device) if getattr(cfg, 'hp_ns_residual', True) else None
        ops.ns_tokenizer_fwd(ns_x, tok.ns_kernel.detach(), tok.ns_bias.detach(), x_ns, 0, C_, L_ns, d, x_hp)
        x, x_hp = engine.candidates_forward(x_ns, C_, self._block_bundles(), cache['plan'], cache['layers'], cfg.num_heads, L_ns,
                                            self.blocks[0].norm1.eps, x_hp)
        return self._heads(x_hp[-C_:] if x_hp is not None else x[-C_:], return_logits)

    def reset_kv_cache(self):
        """OT/model.py:395-397."""
        self.kv_cache = None

    def get_model_info(self) -> Dict:
        """OT/model.py:399-408."""
        return {'total_parameters': sum(p.numel() for p in self.parameters() if p.requires_grad),
                'num_layers': self.config.num_layers, 'hidden_dim': self.config.hidden_dim,
                'num_heads': self.config.num_heads}


def create_onetrans_model(model_type: str = 'default') -> OneTransModel:
    """OT/model.py:411-416."""
    return OneTransModel(get_model_config(model_type))
