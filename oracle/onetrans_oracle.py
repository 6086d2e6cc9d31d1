"""CPU oracle for the OneTrans hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module, and only as the checker / the timed CPU baseline.  Nothing under
``recommend_b200/`` imports it.

What it is: a plain PyTorch (CPU, fp32 or fp64) restatement of the reference's OneTrans model,
``/root/reference/rank/scaling_up/oneTrans/practice/model.py`` ("OT/model.py" below), following the
TensorFlow/Keras 2.12 default semantics the code relies on and adopting the minimal
repairs of SURVEY.md §A.3 where the reference as written cannot run (D2 pyramid indices, D4 weight
alignment, D5 python ints, D6 KV cache, D9 float cast).  Every function cites the reference lines it
follows.

PARITY PIN: the reference ships no tests, golden vectors or seeds and its arithmetic lives in TensorFlow 2.12,
which is absent here (F2).  The oracle is pinned by golden vectors produced by the reference's OWN code: tests/golden/
make_reference_golden.py imports OT/config.py, OT/model.py and OT/data_loader.py UNMODIFIED from /root/reference, with
oracle/tf_shim.py standing in for the ~30 TensorFlow ops they call, runs ``OneTransModel.call`` (pyramid off, 2 blocks; pyramid on,
1 block; a missing sequence), ``OneTransBlock.call`` with its returned (k, v), ``PyramidScheduler.get_layer_config``,
``SequenceProcessor.process_sequence`` and the config classes, and commits inputs, every weight and the outputs
(tests/golden/reference_golden.{npz,json}).  tests/test_reference_golden.py holds this oracle to them at 1e-12 (fp64) in its literal
modes (``ns_param_alignment='head_literal'``, ``query_mode='literal_gather'``, per-token loop) - so the control flow that decides
results is the reference's, executed; only what each TensorFlow op computes is restated (in the shim, from the documented TF 2.12
semantics of SURVEY.md §A.2).  Also from the reference's code: BASELINE config 1 at full size (fp32 oracle within 5e-6 of its
probabilities), gradients of the summed BCE through its forward graph (every parameter tensor, 1e-12), parameter counts, the Keras
weight order.  The same run reproduces defects D2, D6, D7, D8 and D9 as the exceptions the reference raises.  The repaired
modes the product uses are tied to the literal ones by the algebraic invariants in tests/test_oracle.py (T4 tail-only ==
compute-all-then-gather, T5 grouped == per-token loop, T6 causality, hand-computed RMSNorm / BCE vectors) and by the golden vectors
this file generated itself (tests/golden/oracle_golden.json, script tests/golden/make_golden.py).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

# ---------------------------------------------------------------------------------------------------
# configuration  (OT/config.py:9-117)
# ---------------------------------------------------------------------------------------------------


@dataclass
class OracleConfig:
    """Mirror of ``OneTransConfig`` (OT/config.py:12-69), only the fields the model reads."""

    hidden_dim: int = 384          # OT/config.py:14
    num_layers: int = 8            # :15
    num_heads: int = 4             # :16
    ffn_dim: int = 1536            # :17
    num_ns_tokens: int = 12        # :21
    pyramid_enabled: bool = True   # :29
    pyramid_ratios: List[float] = field(
        default_factory=lambda: [0.5, 0.3, 0.2, 0.1, 0.05, 0.03, 0.02, 0.01])  # :30
    dropout_rate: float = 0.1      # :50
    user_features: List[str] = field(default_factory=lambda: ['user_id', 'age', 'gender', 'location'])   # :56
    item_features: List[str] = field(default_factory=lambda: ['item_id', 'category', 'price', 'brand'])  # :57
    context_features: List[str] = field(default_factory=lambda: ['time', 'device', 'platform'])         # :58
    sequence_features: List[str] = field(default_factory=lambda: ['click_seq', 'cart_seq', 'purchase_seq'])  # :59
    tasks: List[str] = field(default_factory=lambda: ['ctr', 'cvr'])  # :63
    seq_feature_dim: int = 64      # OT/model.py:433-442, OT/data_loader.py:322
    # --- additive switches ---
    ns_param_alignment: str = 'tail'          # 'tail' (repair D4) | 'head_literal' (OT/model.py:69-74 as written)
    pyramid_keep_lens: Optional[List[int]] = None  # explicit schedule; None -> reference ratios (repair D2)
    ns_feature_names: Optional[List[str]] = None   # features the NS Dense was built with; None -> all configured

    @property
    def ns_features(self) -> List[str]:
        if self.ns_feature_names is not None:
            return list(self.ns_feature_names)
        return self.user_features + self.item_features + self.context_features  # OT/model.py:243-245


def small_config(**kw) -> OracleConfig:
    """``OneTransSmallConfig`` (OT/config.py:85-92) = the paper's OneTrans-S."""
    return OracleConfig(hidden_dim=256, num_layers=6, ffn_dim=1024, **kw)


def default_config(**kw) -> OracleConfig:
    """``OneTransConfig`` default (OT/config.py:14-17) = the paper's OneTrans-L."""
    return OracleConfig(**kw)


# ---------------------------------------------------------------------------------------------------
# pyramid schedule  (OT/model.py:280-302)
# ---------------------------------------------------------------------------------------------------


def reference_keep_len(layer_idx: int, total_seq_len: int, ratios: Sequence[float]) -> Optional[int]:
    """``keep_len = max(1, int(total_seq_len * keep_ratio))`` in Python double (OT/model.py:292-293).
    Returns None where the reference returns ``query_indices=None`` (OT/model.py:289-290)."""
    if layer_idx >= len(ratios):
        return None
    return max(1, int(total_seq_len * ratios[layer_idx]))


def reference_query_indices(layer_idx: int, total_seq_len: int, ratios: Sequence[float]) -> Optional[List[int]]:
    """Literal ``query_indices`` of OT/model.py:296 (indices into a length-``total_seq_len`` sequence)."""
    k = reference_keep_len(layer_idx, total_seq_len, ratios)
    if k is None:
        return None
    return list(range(total_seq_len - k, total_seq_len))


def keep_lens_reference_ratio(L0: int, num_layers: int, ratios: Sequence[float]) -> List[int]:
    """Per-layer kept tail lengths with repair D2: the ratio applies to the ORIGINAL length L0
    (OT/model.py:343,349 pass the layer-0 length) and the tail is taken from the CURRENT sequence, so
    ``keep = min(keep, cur)``.  Layers past the ratio list keep everything (OT/model.py:289-290)."""
    out, cur = [], L0
    for l in range(num_layers):
        k = reference_keep_len(l, L0, ratios)
        k = cur if k is None else min(k, cur)
        out.append(k)
        cur = k
    return out


def keep_lens_linear_to_ns(L0: int, num_layers: int, L_ns: int) -> List[int]:
    """``L_NS + ((N-1-l)(L0-L_NS))//N`` — "pyramid pruning down to NS tokens" (BASELINE.json config 2;
    PAPER:188).  SURVEY.md §8d."""
    return [L_ns + ((num_layers - 1 - l) * (L0 - L_ns)) // num_layers for l in range(num_layers)]


def keep_lens_halving(L0: int, num_layers: int, L_ns: int) -> List[int]:
    """Query set halved per block, floored at L_NS (BASELINE.json config 4)."""
    out, cur = [], L0
    for _ in range(num_layers):
        cur = max(L_ns, cur // 2)
        out.append(cur)
    return out


def resolve_keep_lens(cfg: OracleConfig, L0: int) -> List[int]:
    if not cfg.pyramid_enabled:
        return [L0] * cfg.num_layers
    if cfg.pyramid_keep_lens is not None:
        assert len(cfg.pyramid_keep_lens) == cfg.num_layers
        out, cur = [], L0
        for k in cfg.pyramid_keep_lens:
            k = max(1, min(int(k), cur))
            out.append(k)
            cur = k
        return out
    return keep_lens_reference_ratio(L0, cfg.num_layers, cfg.pyramid_ratios)


# ---------------------------------------------------------------------------------------------------
# parameters  (Keras kernels are [in, out])
# ---------------------------------------------------------------------------------------------------


def _glorot(shape_in: int, shape_out: int, gen: torch.Generator, lead: Tuple[int, ...] = ()) -> torch.Tensor:
    """Keras Dense default ``glorot_uniform``: U(+-sqrt(6/(fan_in+fan_out)))."""
    lim = math.sqrt(6.0 / (shape_in + shape_out))
    return (torch.rand(*lead, shape_in, shape_out, generator=gen, dtype=torch.float64) * 2 - 1) * lim


def init_params(cfg: OracleConfig, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Random-init parameter dict with Keras default initialisers.  Packed layout:
    index 0 of the leading dimension = shared (S-token) weights, 1+j = dedicated weights of NS token j
    (OT/model.py:38-54, 136-147)."""
    g = torch.Generator().manual_seed(seed)
    d, F, G = cfg.hidden_dim, cfg.ffn_dim, 1 + cfg.num_ns_tokens
    P: Dict[str, torch.Tensor] = {}
    n_ns_feat = len(cfg.ns_features)
    P['tokenizer.ns_tokenizer.kernel'] = _glorot(n_ns_feat, d * cfg.num_ns_tokens, g)       # OT/model.py:212
    P['tokenizer.ns_tokenizer.bias'] = torch.zeros(d * cfg.num_ns_tokens, dtype=torch.float64)
    for i in range(len(cfg.sequence_features)):                                              # OT/model.py:217-219
        P[f'tokenizer.seq_projections.{i}.kernel'] = _glorot(cfg.seq_feature_dim, d, g)
        P[f'tokenizer.seq_projections.{i}.bias'] = torch.zeros(d, dtype=torch.float64)
    P['tokenizer.sep_embedding'] = (torch.rand(1, d, generator=g, dtype=torch.float64) - 0.5) * 0.1  # :222, U(-.05,.05)
    for l in range(cfg.num_layers):
        b = f'blocks.{l}.'
        P[b + 'norm1.scale'] = torch.ones(d, dtype=torch.float64)                            # OT/model.py:16
        P[b + 'norm2.scale'] = torch.ones(d, dtype=torch.float64)
        P[b + 'attention.Wq'] = _glorot(d, d, g, (G,))                                        # :38,43-46
        P[b + 'attention.Wk'] = _glorot(d, d, g, (G,))                                        # :39,47-50
        P[b + 'attention.Wv'] = _glorot(d, d, g, (G,))                                        # :40,51-54
        P[b + 'attention.Wo'] = _glorot(d, d, g)                                              # :57
        P[b + 'ffn.W1'] = _glorot(d, F, g, (G,))                                              # :137,144
        P[b + 'ffn.b1'] = torch.zeros(G, F, dtype=torch.float64)
        P[b + 'ffn.W2'] = _glorot(F, d, g, (G,))                                              # :138,145
        P[b + 'ffn.b2'] = torch.zeros(G, d, dtype=torch.float64)
    P['output_norm.scale'] = torch.ones(d, dtype=torch.float64)                               # :322
    for t in cfg.tasks:                                                                       # :325-330
        P[f'task_heads.{t}.0.kernel'] = _glorot(d, d // 2, g)
        P[f'task_heads.{t}.0.bias'] = torch.zeros(d // 2, dtype=torch.float64)
        P[f'task_heads.{t}.1.kernel'] = _glorot(d // 2, 1, g)
        P[f'task_heads.{t}.1.bias'] = torch.zeros(1, dtype=torch.float64)
    return {k: v.to(dtype) for k, v in P.items()}


def randomize_small_params(P: Dict[str, torch.Tensor], seed: int = 1, scale: float = 0.1) -> None:
    """Perturb biases / gains / SEP away from their (zero / one) defaults so that parity tests exercise
    them.  In place."""
    g = torch.Generator().manual_seed(seed)
    for k, v in P.items():
        if k.endswith('.bias') or k.endswith('.b1') or k.endswith('.b2'):
            v.copy_((torch.rand(v.shape, generator=g, dtype=torch.float64) * 2 - 1).to(v.dtype) * scale)
        elif k.endswith('.scale'):
            v.copy_(1.0 + (torch.rand(v.shape, generator=g, dtype=torch.float64) * 2 - 1).to(v.dtype) * scale)


def count_params(P: Dict[str, torch.Tensor]) -> int:
    return sum(v.numel() for v in P.values())


# ---------------------------------------------------------------------------------------------------
# primitives
# ---------------------------------------------------------------------------------------------------


def rmsnorm(x: torch.Tensor, scale: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """OT/model.py:19-23: ``x * rsqrt(mean(x^2, -1) + eps) * scale``."""
    variance = torch.mean(torch.square(x), dim=-1, keepdim=True)
    x = x * torch.rsqrt(variance + eps)
    return x * scale


def gelu_erf(x: torch.Tensor) -> torch.Tensor:
    """Keras ``activation='gelu'`` = exact erf form."""
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def group_of_position(p: int, cur: int, L_ns: int, alignment: str) -> int:
    """Weight group (0 = shared, 1+j = dedicated j) of position ``p`` in a length-``cur`` sequence.
    'head_literal' is OT/model.py:69-74 as written (``idx < num_ns_tokens`` -> dedicated[idx]);
    'tail' is repair D4: the NS tokens sit at the tail (OT/model.py:235), token j of the ORIGINAL
    L_ns NS tokens keeps dedicated[j] however much of the sequence in front of it was pruned."""
    if alignment == 'head_literal':
        return 1 + p if p < L_ns else 0
    if alignment == 'tail':
        return 1 + (L_ns - (cur - p)) if p >= cur - L_ns else 0
    raise ValueError(alignment)


def _mixed_linear(x: torch.Tensor, W: torch.Tensor, b: Optional[torch.Tensor], first_pos: int, cur: int,
                  L_ns: int, alignment: str, literal_loop: bool) -> torch.Tensor:
    """Apply the per-position Dense of OT/model.py:84-88 / 154-161 to rows ``first_pos..`` of a
    length-``cur`` sequence.  x: [B, n, in]; W: [G, in, out]; b: [G, out] or None."""
    n = x.shape[1]
    groups = [group_of_position(first_pos + i, cur, L_ns, alignment) for i in range(n)]
    if literal_loop:  # the reference's dispatch structure: one tiny matmul per position
        outs = []
        for i, gi in enumerate(groups):
            y = x[:, i:i + 1, :] @ W[gi]
            if b is not None:
                y = y + b[gi]
            outs.append(y)
        return torch.cat(outs, dim=1)
    gidx = torch.tensor(groups, dtype=torch.long)
    out = torch.empty(x.shape[0], n, W.shape[2], dtype=x.dtype)
    shared = gidx == 0
    if shared.any():
        y = x[:, shared, :] @ W[0]
        if b is not None:
            y = y + b[0]
        out[:, shared, :] = y
    ded = (~shared).nonzero().flatten().tolist()
    if ded:
        Wd = W[gidx[ded]]                                  # [n_d, in, out]
        y = torch.einsum('bni,nio->bno', x[:, ded, :], Wd)
        if b is not None:
            y = y + b[gidx[ded]]
        out[:, ded, :] = y
    return out


def mixed_mha(P: Dict[str, torch.Tensor], prefix: str, cfg: OracleConfig, x: torch.Tensor, keep: int,
              literal_loop: bool = False, return_kv: bool = False):
    """OT/model.py:76-122 on a normalised input ``x`` [B, cur, d]; queries only for the last ``keep``
    rows (D3: the reference computes every query then discards, OT/model.py:356-371; same result).
    Scores scaled by 1/sqrt(dh) (OT/model.py:106), causal lower-triangular mask with the query tail
    aligned to the key tail (allowed iff k <= q + (cur-keep)), masked entries REPLACED by -1e9
    (OT/model.py:109-110), softmax over keys (:112), Wo without bias (:117)."""
    B, cur, d = x.shape
    H, dh, L_ns = cfg.num_heads, cfg.hidden_dim // cfg.num_heads, cfg.num_ns_tokens
    al = cfg.ns_param_alignment
    k = _mixed_linear(x, P[prefix + 'Wk'], None, 0, cur, L_ns, al, literal_loop)
    v = _mixed_linear(x, P[prefix + 'Wv'], None, 0, cur, L_ns, al, literal_loop)
    q = _mixed_linear(x[:, cur - keep:, :], P[prefix + 'Wq'], None, cur - keep, cur, L_ns, al, literal_loop)
    q = q.reshape(B, keep, H, dh)
    k4 = k.reshape(B, cur, H, dh)
    v4 = v.reshape(B, cur, H, dh)
    scores = torch.einsum('bqhd,bkhd->bhqk', q, k4) / math.sqrt(float(dh))
    qi = torch.arange(keep).unsqueeze(1) + (cur - keep)
    ki = torch.arange(cur).unsqueeze(0)
    allowed = ki <= qi
    scores = torch.where(allowed, scores, torch.full_like(scores, -1e9))
    w = torch.softmax(scores, dim=-1)
    o = torch.einsum('bhqk,bkhd->bqhd', w, v4).reshape(B, keep, d)
    out = o @ P[prefix + 'Wo']
    if return_kv:
        return out, (k, v)
    return out


def mixed_ffn(P: Dict[str, torch.Tensor], prefix: str, cfg: OracleConfig, x: torch.Tensor, first_pos: int, cur: int,
              literal_loop: bool = False) -> torch.Tensor:
    """OT/model.py:149-163: ``Dense(F, gelu) -> Dense(d)`` with the position's weight group."""
    L_ns, al = cfg.num_ns_tokens, cfg.ns_param_alignment
    h = gelu_erf(_mixed_linear(x, P[prefix + 'W1'], P[prefix + 'b1'], first_pos, cur, L_ns, al, literal_loop))
    return _mixed_linear(h, P[prefix + 'W2'], P[prefix + 'b2'], first_pos, cur, L_ns, al, literal_loop)


def _dropout(x: torch.Tensor, rate: float, training: bool, gen: Optional[torch.Generator]) -> torch.Tensor:
    """Keras inverted dropout; identity unless training."""
    if not training or rate <= 0.0:
        return x
    keep = (torch.rand(x.shape, generator=gen, dtype=torch.float32) >= rate).to(x.dtype)
    return x * keep / (1.0 - rate)


def block_forward(P, l: int, cfg: OracleConfig, x: torch.Tensor, keep: int, query_mode: str = 'tail_only',
                  literal_loop: bool = False, training: bool = False, gen=None) -> torch.Tensor:
    """OneTransBlock.call (OT/model.py:186-200) followed by the tail gather of OT/model.py:371.
    query_mode 'tail_only' computes only the kept rows; 'literal_gather' runs the block on every row
    and gathers afterwards exactly as the reference does."""
    b = f'blocks.{l}.'
    cur = x.shape[1]
    if query_mode == 'literal_gather':
        xn = rmsnorm(x, P[b + 'norm1.scale'])
        a = mixed_mha(P, b + 'attention.', cfg, xn, cur, literal_loop)
        z = x + _dropout(a, cfg.dropout_rate, training, gen)
        zn = rmsnorm(z, P[b + 'norm2.scale'])
        f = mixed_ffn(P, b + 'ffn.', cfg, zn, 0, cur, literal_loop)
        y = z + _dropout(f, cfg.dropout_rate, training, gen)
        return y[:, cur - keep:, :]           # tf.gather(block_output, query_indices, axis=1), OT/model.py:371
    xn = rmsnorm(x, P[b + 'norm1.scale'])                                  # :191
    a = mixed_mha(P, b + 'attention.', cfg, xn, keep, literal_loop)         # :192
    z = x[:, cur - keep:, :] + _dropout(a, cfg.dropout_rate, training, gen)  # :193
    zn = rmsnorm(z, P[b + 'norm2.scale'])                                  # :196
    f = mixed_ffn(P, b + 'ffn.', cfg, zn, cur - keep, cur, literal_loop)    # :197
    return z + _dropout(f, cfg.dropout_rate, training, gen)                # :198


def tokenizer_forward(P, cfg: OracleConfig, non_seq: Dict[str, torch.Tensor], seq: Dict[str, torch.Tensor]) -> torch.Tensor:
    """Tokenizer.call (OT/model.py:224-277): S tokens first, NS tokens last (:235)."""
    dt = P['tokenizer.sep_embedding'].dtype
    d, L_ns = cfg.hidden_dim, cfg.num_ns_tokens
    # ---- non-sequence features (OT/model.py:239-254); every feature cast to float (repair D9) ----
    feats = [non_seq[n].to(dt).reshape(-1, 1) for n in cfg.ns_features if n in non_seq]
    ref = next(iter(non_seq.values())) if non_seq else next(iter(seq.values()))
    B = ref.shape[0]
    if not feats:
        ns = torch.zeros(B, L_ns, d, dtype=dt)                              # :249-251
    else:
        cat = torch.cat(feats, dim=-1)                                      # :253
        ns = (cat @ P['tokenizer.ns_tokenizer.kernel'] + P['tokenizer.ns_tokenizer.bias']).reshape(B, L_ns, d)  # :211-214
    # ---- sequence features (OT/model.py:256-277) ----
    toks = []
    n_seq = len(cfg.sequence_features)
    for i, name in enumerate(cfg.sequence_features):
        if name in seq:
            e = seq[name].to(dt)
            toks.append(e @ P[f'tokenizer.seq_projections.{i}.kernel'] + P[f'tokenizer.seq_projections.{i}.bias'])  # :265
            if i < n_seq - 1:                                               # :269
                toks.append(P['tokenizer.sep_embedding'][0].expand(B, 1, d))  # :270-272
    s = torch.cat(toks, dim=1) if toks else torch.zeros(B, 0, d, dtype=dt)  # :274-277
    return torch.cat([s, ns], dim=1)                                        # :235


def heads_forward(P, cfg: OracleConfig, x_last: torch.Tensor) -> Dict[str, torch.Tensor]:
    """Task heads on the last token (OT/model.py:388-391): Dense(d/2, gelu) -> Dense(1).  Returns
    LOGITS (pre-sigmoid, D17); probabilities = sigmoid(logits) (OT/model.py:329)."""
    out = {}
    for t in cfg.tasks:
        h = gelu_erf(x_last @ P[f'task_heads.{t}.0.kernel'] + P[f'task_heads.{t}.0.bias'])
        out[t] = h @ P[f'task_heads.{t}.1.kernel'] + P[f'task_heads.{t}.1.bias']
    return out


def model_forward(P, cfg: OracleConfig, non_seq, seq, training: bool = False, query_mode: str = 'tail_only',
                  literal_loop: bool = False, return_logits: bool = False, gen=None,
                  return_hidden: bool = False):
    """OneTransModel.call (OT/model.py:335-393)."""
    x = tokenizer_forward(P, cfg, non_seq, seq)                              # :342
    keep_lens = resolve_keep_lens(cfg, x.shape[1])                           # :349 + repairs D2/D5
    hidden = [x]
    for l in range(cfg.num_layers):                                          # :348
        x = block_forward(P, l, cfg, x, keep_lens[l], query_mode, literal_loop, training, gen)
        hidden.append(x)
    xo = rmsnorm(x, P['output_norm.scale'])                                  # :384
    logits = heads_forward(P, cfg, xo[:, -1, :])                             # :390
    out = logits if return_logits else {t: torch.sigmoid(v) for t, v in logits.items()}
    if return_hidden:
        return out, hidden
    return out


def bce_loss(probs: Dict[str, torch.Tensor], labels: Dict[str, torch.Tensor], tasks: Sequence[str]) -> torch.Tensor:
    """Sum over tasks of Keras ``BinaryCrossentropy(from_logits=False)`` (OT/train.py:84-87,124-128):
    clip p to [1e-7, 1-1e-7], ``-[y log(p+1e-7) + (1-y) log(1-p+1e-7)]``, mean over batch."""
    eps = 1e-7
    total = 0.0
    for t in tasks:
        if t in probs and t in labels:
            p = torch.clamp(probs[t], eps, 1.0 - eps)
            y = labels[t].to(p.dtype)
            bce = -(y * torch.log(p + eps) + (1.0 - y) * torch.log(1.0 - p + eps))
            total = total + bce.mean(dim=-1).mean()
    return total


def loss_and_grads(P, cfg: OracleConfig, non_seq, seq, labels, **kw):
    """Forward + BCE + ``tape.gradient`` (OT/train.py:116-131) via torch autograd on the oracle."""
    Pg = {k: v.detach().clone().requires_grad_(True) for k, v in P.items()}
    probs = model_forward(Pg, cfg, non_seq, seq, **kw)
    loss = bce_loss(probs, labels, cfg.tasks)
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in Pg.items()}
    return loss.detach(), grads, {t: v.detach() for t, v in probs.items()}


def clip_by_norm(g: torch.Tensor, c: float) -> torch.Tensor:
    """``tf.clip_by_norm`` per tensor (OT/train.py:135): g * c / max(||g||, c)."""
    n = torch.linalg.vector_norm(g)
    return g * (c / torch.maximum(n, torch.tensor(c, dtype=g.dtype)))


def rmsprop_step(w: torch.Tensor, g: torch.Tensor, rms: torch.Tensor, mom: Optional[torch.Tensor], lr: float,
                 rho: float = 0.9, momentum: float = 0.0, eps: float = 1e-7):
    """One Keras-2.12 ``RMSprop.update_step`` (optimizer built at OT/train.py:65-70, applied at :138; the arithmetic
    is TensorFlow's): ``rms = rho*rms + (1-rho)*g^2``; ``inc = lr*g*rsqrt(rms + eps)`` (non-centered:
    epsilon inside the root); with momentum ``mom = momentum*mom + inc; w -= mom`` else ``w -= inc``.
    Returns the new (w, rms, mom); inputs are not modified."""
    rms = rho * rms + (1.0 - rho) * g * g
    inc = lr * g * torch.rsqrt(rms + eps)
    if momentum > 0.0:
        mom = momentum * mom + inc
        return w - mom, rms, mom
    return w - inc, rms, mom


PACKED_OVER_GROUPS = ('.attention.Wq', '.attention.Wk', '.attention.Wv', '.ffn.W1', '.ffn.b1', '.ffn.W2', '.ffn.b2')


def clip_per_keras_variable(name, g: torch.Tensor, c: float) -> torch.Tensor:
    """``[tf.clip_by_norm(g, c) for g in gradients]`` (OT/train.py:133-135) runs over ``model.trainable_variables``, i.e. per
    KERAS variable.  The oracle packs the per-position Dense layers of OT/model.py:38-54 / 136-147 over a leading weight-group
    index (0 = shared, 1+j = dedicated j): each index of those tensors is a Keras variable of its own and is clipped on its own."""
    if isinstance(name, str) and name.endswith(PACKED_OVER_GROUPS):
        return torch.stack([clip_by_norm(g[i], c) for i in range(g.shape[0])])
    return clip_by_norm(g, c)


def clip_rmsprop_update(params: Dict[str, torch.Tensor], grads: Dict[str, torch.Tensor], state: Dict[str, Dict[str, torch.Tensor]],
                        lr: float = 0.005, rho: float = 0.9, momentum: float = 0.99999, eps: float = 1e-7,
                        clip_norm: float = 90.0) -> None:
    """OT/train.py:133-138 for every trainable variable: ``tf.clip_by_norm`` per Keras variable (see
    ``clip_per_keras_variable``), then ``apply_gradients``.
    Defaults are OT/config.py:39-52 (dense_lr, momentum, gradient_clip_norm).  Updates ``params``/``state`` in place."""
    for name, w in params.items():
        g = grads[name]
        if clip_norm > 0:
            g = clip_per_keras_variable(name, g, clip_norm)
        st = state.setdefault(name, {'rms': torch.zeros_like(w), 'mom': torch.zeros_like(w)})
        nw, st['rms'], nm = rmsprop_step(w, g, st['rms'], st['mom'], lr, rho, momentum, eps)
        if nm is not None:
            st['mom'] = nm
        params[name] = nw


def embed_events(table: torch.Tensor, vocab_sizes: Sequence[int], ids: torch.Tensor, out_dtype=torch.bfloat16) -> torch.Tensor:
    """ID front end of the sequence tokenizer (extension along PAPER:89-109; lookup-and-concat idiom of
    recall/bert_like/kuaiformer/practice/model.py:58-94): ``ids [..., n_fields]`` -> concat over fields of
    ``table[field_off[f] + ids[..., f]]``, rounded to the dtype the tokenizer reads its events in."""
    off, cols = 0, []
    for f, v in enumerate(vocab_sizes):
        cols.append(table[off + ids[..., f].long()])
        off += v
    return torch.cat(cols, dim=-1).to(out_dtype)


def embed_grad_table(table: torch.Tensor, vocab_sizes: Sequence[int], ids: torch.Tensor, d_events: torch.Tensor) -> torch.Tensor:
    """Dense view of the sparse gradient ``tape.gradient`` gives an Embedding (IndexedSlices summed per row)."""
    g = torch.zeros_like(table, dtype=torch.float64)
    ef = table.shape[1]
    off = 0
    flat_ids = ids.reshape(-1, len(vocab_sizes))
    d2 = d_events.reshape(-1, len(vocab_sizes) * ef).double()
    for f, v in enumerate(vocab_sizes):
        g.index_add_(0, off + flat_ids[:, f].long(), d2[:, f * ef:(f + 1) * ef])
        off += v
    return g


def adagrad_step(w: torch.Tensor, g: torch.Tensor, acc: torch.Tensor, lr: float = 0.1, eps: float = 1e-7):
    """Keras-2.12 ``Adagrad.update_step`` (sparse_optimizer of OT/config.py:39-47; initial accumulator 0.1):
    ``acc += g^2; w -= lr * g / sqrt(acc + eps)`` - epsilon inside the root (the legacy TF1 op put it outside).
    Returns the new (w, acc)."""
    acc = acc + g * g
    return w - lr * g / torch.sqrt(acc + eps), acc


# ---------------------------------------------------------------------------------------------------
# synthetic inputs  (OT/data_loader.py:301-329, :146-154)
# ---------------------------------------------------------------------------------------------------


def synthetic_batch(cfg: OracleConfig, B: int, seq_lens: Sequence[int], seed: int = 1234, ns_mode: str = 'normal',
                    dtype=torch.float32):
    """Seeded synthetic batch.  ns_mode 'ids' follows create_sample_batch (ids randint(0,100)/(0,1000)
    cast to float, context U[0,1)); 'normal' uses N(0,1) (throughput / bf16 runs)."""
    g = torch.Generator().manual_seed(seed)
    non_seq = {}
    if ns_mode == 'ids':
        for n in cfg.user_features:
            non_seq[n] = torch.randint(0, 100, (B, 1), generator=g).to(dtype)       # OT/data_loader.py:310
        for n in cfg.item_features:
            non_seq[n] = torch.randint(0, 1000, (B, 1), generator=g).to(dtype)      # :313
        for n in cfg.context_features:
            non_seq[n] = torch.rand(B, 1, generator=g, dtype=torch.float64).to(dtype)  # :316
    else:
        for n in cfg.ns_features:
            non_seq[n] = torch.randn(B, 1, generator=g, dtype=torch.float64).to(dtype)
    seq = {}
    for name, L in zip(cfg.sequence_features, seq_lens):
        seq[name] = torch.randn(B, L, cfg.seq_feature_dim, generator=g, dtype=torch.float64).to(dtype)  # OT/data_loader.py:146
    labels = {t: (torch.rand(B, 1, generator=g) < 0.5).to(dtype) for t in cfg.tasks}  # OT/data_loader.py:151-154
    return non_seq, seq, labels


# ---------------------------------------------------------------------------------------------------
# two-stage inference with a per-layer cache of the sequence-side K/V  (PAPER:144-151; the reference's own
# branch OT/model.py:95-98,120 cannot run).  Pinned in tests/test_oracle.py: stage 1 + stage 2 equal
# ``model_forward`` on the same rows (S rows never see NS rows: causal mask, S first, OT/model.py:109-110, 235), and
# with the pyramid off ``two_stage_extend`` equals a fresh stage 1 on the longer sequence.
# ---------------------------------------------------------------------------------------------------
def _layer_plan(L0: int, L_ns: int, keep_lens: Sequence[int]):
    """Per layer (cur, Tn, cur_S, keep, Tq, keep_S): length, NS tokens alive, S tokens alive, kept tail, NS / S tokens that query."""
    plan, cur = [], L0
    for keep in keep_lens:
        Tn = min(L_ns, cur)
        Tq = min(Tn, keep)
        plan.append((cur, Tn, cur - Tn, keep, Tq, keep - Tq))
        cur = keep
    return plan


def _attend_tail(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, H: int) -> torch.Tensor:
    """q [B, Tq, d] are the LAST Tq positions of the key sequence k/v [B, Lk, d]: query i sees keys 0 .. (Lk - Tq) + i
    (OT/model.py:101-114 with the tail-aligned mask of ``mixed_mha``)."""
    B, Tq, d = q.shape
    Lk, dh = k.shape[1], d // H
    s = torch.einsum('bqhd,bkhd->bhqk', q.reshape(B, Tq, H, dh), k.reshape(B, Lk, H, dh)) / math.sqrt(float(dh))
    allowed = torch.arange(Lk).unsqueeze(0) <= torch.arange(Tq).unsqueeze(1) + (Lk - Tq)
    s = torch.where(allowed, s, torch.full_like(s, -1e9))
    return torch.einsum('bhqk,bkhd->bqhd', torch.softmax(s, dim=-1), v.reshape(B, Lk, H, dh)).reshape(B, Tq, d)


def _s_rows_through_block(P, l: int, cfg: OracleConfig, x: torch.Tensor, n_q: int, kv_old=None):
    """Sequence tokens (shared weights, group 0) through block ``l``: K/V for every row of ``x`` (behind ``kv_old`` if given),
    queries for the last ``n_q`` rows.  Returns ``(y [1, n_q, d], (k, v))``."""
    b = f'blocks.{l}.'
    xn = rmsnorm(x, P[b + 'norm1.scale'])
    k, v = xn @ P[b + 'attention.Wk'][0], xn @ P[b + 'attention.Wv'][0]
    if kv_old is not None:
        k, v = torch.cat([kv_old[0], k], 1), torch.cat([kv_old[1], v], 1)
    if n_q == 0:
        return x[:, :0], (k, v)
    q = xn[:, xn.shape[1] - n_q:] @ P[b + 'attention.Wq'][0]
    z = x[:, x.shape[1] - n_q:] + _attend_tail(q, k, v, cfg.num_heads) @ P[b + 'attention.Wo']
    zn = rmsnorm(z, P[b + 'norm2.scale'])
    f = gelu_erf(zn @ P[b + 'ffn.W1'][0] + P[b + 'ffn.b1'][0]) @ P[b + 'ffn.W2'][0] + P[b + 'ffn.b2'][0]
    return z + f, (k, v)


def two_stage_user_cache(P, cfg: OracleConfig, seq: Dict[str, torch.Tensor]):
    """Stage 1, one user (batch 1): every layer's sequence-side (K, V)."""
    assert cfg.ns_param_alignment == 'tail'
    x = tokenizer_forward(P, cfg, {}, seq)
    x = x[:, :x.shape[1] - cfg.num_ns_tokens]
    L0 = x.shape[1] + cfg.num_ns_tokens
    plan = _layer_plan(L0, cfg.num_ns_tokens, resolve_keep_lens(cfg, L0))
    layers = []
    for l, (cur, Tn, cur_S, keep, Tq, keep_S) in enumerate(plan):
        if cur_S == 0:
            layers.append(None)
            x = x[:, :0]
            continue
        assert x.shape[1] == cur_S
        x, kv = _s_rows_through_block(P, l, cfg, x, keep_S)
        layers.append(kv)
    return {'plan': plan, 'layers': layers, 'L0': L0}


def two_stage_extend(P, cfg: OracleConfig, cache, new_events: torch.Tensor):
    """Cross-request incremental update (PAPER:151): ``new_events [1, n, 64]`` are new behaviours of the LAST sequence, i.e. new
    tokens at the tail of the S block.  Streaming rule: layer l gains the K/V of the new tokens alive at l; a new token sees
    every key the layer already holds plus the new ones up to itself; ``n_{l+1} = min(n_l, keep_S(l))`` of them move on (the
    key sets of a built cache only grow - which old tokens a layer keeps was decided when the cache was built)."""
    i = len(cfg.sequence_features) - 1
    x = new_events.to(P['tokenizer.sep_embedding'].dtype) @ P[f'tokenizer.seq_projections.{i}.kernel'] + P[f'tokenizer.seq_projections.{i}.bias']
    layers = list(cache['layers'])
    for l, (cur, Tn, cur_S, keep, Tq, keep_S) in enumerate(cache['plan']):
        if layers[l] is None or x.shape[1] == 0:
            break
        x, layers[l] = _s_rows_through_block(P, l, cfg, x, min(x.shape[1], keep_S), layers[l])
    return {'plan': cache['plan'], 'layers': layers, 'L0': cache['L0']}


def two_stage_score(P, cfg: OracleConfig, cache, non_seq: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Stage 2: logits of C candidates of the cached user; only their NS tokens are computed."""
    L_ns, H = cfg.num_ns_tokens, cfg.num_heads
    x = tokenizer_forward(P, cfg, non_seq, {})                       # [C, L_ns, d]: no sequence -> NS tokens only
    for l, ((cur, Tn, cur_S, keep, Tq, keep_S), kv_s) in enumerate(zip(cache['plan'], cache['layers'])):
        b = f'blocks.{l}.'
        C = x.shape[0]
        assert x.shape[1] == Tn
        xn = rmsnorm(x, P[b + 'norm1.scale'])
        lin = lambda t, W, bias, first: _mixed_linear(t, P[b + W], None if bias is None else P[b + bias], first, cur, L_ns, 'tail', False)
        k, v = lin(xn, 'attention.Wk', None, cur - Tn), lin(xn, 'attention.Wv', None, cur - Tn)
        if kv_s is not None:
            k = torch.cat([kv_s[0].expand(C, -1, -1), k], 1)
            v = torch.cat([kv_s[1].expand(C, -1, -1), v], 1)
        q = lin(xn[:, Tn - Tq:], 'attention.Wq', None, cur - Tq)
        z = x[:, Tn - Tq:] + _attend_tail(q, k, v, H) @ P[b + 'attention.Wo']
        zn = rmsnorm(z, P[b + 'norm2.scale'])
        x = z + lin(gelu_erf(lin(zn, 'ffn.W1', 'ffn.b1', cur - Tq)), 'ffn.W2', 'ffn.b2', cur - Tq)
    return heads_forward(P, cfg, rmsnorm(x, P['output_norm.scale'])[:, -1, :])
