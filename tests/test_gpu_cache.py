"""T8: scoring C candidates of one user with the cached sequence-side K/V equals the full
uncached forward on the same C rows, and the cached-attention kernel equals a plain fp32 reference."""
import math

import pytest
import torch

from oracle import onetrans_oracle as O
import recommend_b200 as R
from recommend_b200 import ops
from tests.helpers import make_configs, to_cuda, rel_l2

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16
LOGIT_TOL = 1e-2      # north_star: logits rel err <= 1e-2 for the bf16 kernels against the fp32 oracle - cached path included


@pytest.mark.parametrize('C,H,dh,Tq,Tn,Ls', [(40, 4, 64, 32, 32, 300), (7, 4, 64, 16, 16, 100), (33, 4, 64, 12, 12, 0),
                                             (19, 4, 64, 5, 13, 77), (16, 4, 96, 8, 8, 150), (1000, 4, 64, 32, 32, 512)])
def test_cached_attention_kernel(C, H, dh, Tq, Tn, Ls):
    d = H * dh
    g = torch.Generator(device='cuda').manual_seed(0)
    rnd = lambda *s: torch.randn(*s, generator=g, device='cuda').to(bf16)
    q, kv_own, kv_s = rnd(Tq * C, d), rnd(Tn * C, 2 * d), rnd(max(Ls, 1), 2 * d)
    o = torch.full((Tq * C, d), float('nan'), dtype=bf16, device='cuda')
    ops.attn_ns_cached(q, kv_own[:, :d], kv_own[:, d:], kv_s[:, :d] if Ls else None, kv_s[:, d:] if Ls else None, o, C, H, Tq, Tn, Ls, dh)
    # reference: per candidate, keys = [shared ; own], query i sees all shared keys and own keys 0..(Tn-Tq)+i
    q4 = q.float().view(Tq, C, H, dh).permute(1, 2, 0, 3)
    ko = kv_own[:, :d].float().reshape(Tn, C, H, dh).permute(1, 2, 0, 3)
    vo = kv_own[:, d:].float().reshape(Tn, C, H, dh).permute(1, 2, 0, 3)
    ks = kv_s[:Ls, :d].float().reshape(Ls, H, dh).permute(1, 0, 2).unsqueeze(0).expand(C, H, Ls, dh)
    vs = kv_s[:Ls, d:].float().reshape(Ls, H, dh).permute(1, 0, 2).unsqueeze(0).expand(C, H, Ls, dh)
    k, v = torch.cat([ks, ko], 2), torch.cat([vs, vo], 2)
    s = q4 @ k.transpose(-1, -2) / math.sqrt(dh)
    qi = torch.arange(Tq, device='cuda')[:, None] + (Tn - Tq)
    ki = torch.arange(Ls + Tn, device='cuda')[None, :] - Ls
    s = torch.where(ki <= qi, s, torch.full_like(s, -1e9))
    ref = (torch.softmax(s, -1) @ v).permute(2, 0, 1, 3).reshape(Tq * C, d)
    assert not torch.isnan(o.float()).any()
    assert ((o.float() - ref).abs() / (1 + ref.abs())).max().item() < 2e-2


@pytest.mark.parametrize('schedule,L_ns,layers', [('linear_to_ns', 16, 4), ('reference_ratio', 16, 6), ('halving', 8, 3)])
def test_t8_cached_scoring_equals_uncached_forward(schedule, L_ns, layers):
    """T8 against the ORACLE: the cached two-stage scoring and the uncached forward on the same C rows are each held to the
    north_star bar (logits rel-L2 <= 1e-2 vs the fp32 oracle, identical bf16-representable weights).  The two bf16 evaluations are
    also compared with each other, but only against the triangle bound of the two (2e-2): they round at different places (tile
    packing, softmax block order, which norms ride in an epilogue), so their mutual distance is not a parity statement."""
    ocfg, cfg = make_configs(num_layers=layers, num_ns_tokens=L_ns, schedule=schedule)
    P = O.init_params(ocfg, seed=5)
    O.randomize_small_params(P, seed=6)
    Pb = {k: (v.to(bf16).float() if (v.dim() >= 2 and 'ns_tokenizer' not in k and 'task_heads' not in k and 'sep_embedding' not in k) else v)
          for k, v in P.items()}
    model = R.OneTransModel(cfg).cuda()
    R.load_reference_style_params(model, Pb)
    C = 50
    non_seq, seq1, _ = O.synthetic_batch(ocfg, C, (60, 50, 40), seed=11)
    seq1 = {k: v[:1] for k, v in seq1.items()}                      # ONE user
    seqC = {k: v.expand(C, -1, -1).contiguous() for k, v in seq1.items()}
    with torch.no_grad():
        full = model(to_cuda(non_seq), to_cuda(seqC), return_logits=True)
        model.reset_kv_cache()
        cached = model(to_cuda(non_seq), to_cuda(seq1), use_kv_cache=True, return_logits=True)   # builds the cache
        again = model(to_cuda(non_seq), to_cuda(seq1), use_kv_cache=True, return_logits=True)    # reuses it
    seq_o = {k: v.to(bf16).float() for k, v in seqC.items()}
    ocfg.pyramid_keep_lens = R.resolve_keep_lens(cfg, 60 + 50 + 40 + 2 + L_ns)
    ref = O.model_forward(Pb, ocfg, non_seq, seq_o, return_logits=True)
    cat = lambda d: torch.cat([d[t].flatten().float().cpu() for t in cfg.tasks])
    lo = cat(ref)
    e_cached, e_full, e_mutual = rel_l2(cat(cached), lo), rel_l2(cat(full), lo), rel_l2(cat(cached), cat(full))
    print(f'cached vs fp32 oracle rel-L2 {e_cached:.3e}   uncached vs fp32 oracle {e_full:.3e}   cached vs uncached {e_mutual:.3e}')
    for t in cfg.tasks:
        assert cached[t].shape == (C, 1)
        assert torch.equal(cached[t], again[t])
    assert e_cached <= LOGIT_TOL, e_cached              # north_star: logits rel err <= 1e-2, cached path
    assert e_full <= LOGIT_TOL, e_full                  # ... and the uncached path on the same rows
    assert e_mutual <= 2 * LOGIT_TOL, e_mutual


def _cache_fp(cache):
    return [None if kv is None else kv.float().cpu() for kv in cache['layers']]


@pytest.mark.parametrize('schedule,pyramid,L_ns,layers,n_new', [('linear_to_ns', True, 16, 4, 5), ('reference_ratio', True, 16, 6, 40),
                                                              ('halving', True, 8, 3, 1), ('linear_to_ns', False, 8, 3, 7)])
def test_extend_kv_cache_matches_the_oracle_streaming_rule(schedule, pyramid, L_ns, layers, n_new):
    """append n new behaviours to a built cache == oracle ``two_stage_extend`` (fp32) on the same weights."""
    ocfg, cfg = make_configs(num_layers=layers, num_ns_tokens=L_ns, schedule=schedule, pyramid_enabled=pyramid)
    P = O.init_params(ocfg, seed=5)
    O.randomize_small_params(P, seed=6)
    # bf16-representable GEMM weights (as in T8 above): the comparison measures the kernels, not the weight cast
    P = {k: (v.to(bf16).float() if (v.dim() >= 2 and 'ns_tokenizer' not in k and 'task_heads' not in k and 'sep_embedding' not in k) else v)
         for k, v in P.items()}
    model = R.OneTransModel(cfg).cuda()
    R.load_reference_style_params(model, P)
    C = 30
    non_seq, seq, _ = O.synthetic_batch(ocfg, C, (60, 50, 40 + n_new), seed=11)
    seq = {k: v[:1].to(bf16).float() for k, v in seq.items()}                       # one user; bf16-representable events
    last = ocfg.sequence_features[-1]
    head = dict(seq)
    head[last], new = seq[last][:, :40], seq[last][:, 40:]
    L0 = 60 + 50 + 40 + 2 + L_ns
    ocfg.pyramid_keep_lens = R.resolve_keep_lens(cfg, L0) if pyramid else None
    want_cache = O.two_stage_extend(P, ocfg, O.two_stage_user_cache(P, ocfg, head), new)
    want = O.two_stage_score(P, ocfg, want_cache, non_seq)
    with torch.no_grad():
        model.build_kv_cache(to_cuda(head))
        before = _cache_fp(model.kv_cache)
        model.extend_kv_cache(new.cuda())
        got = model.score_candidates(to_cuda(non_seq), return_logits=True)
    d = cfg.hidden_dim
    for l, (kv, want_kv, old) in enumerate(zip(_cache_fp(model.kv_cache), want_cache['layers'], before)):
        if want_kv is None:
            assert kv is None
            continue
        wk, wv = want_kv
        assert kv.shape[0] == wk.shape[1], (l, kv.shape, wk.shape)                  # same key-set sizes per layer
        assert torch.equal(kv[:old.shape[0]], old)                                  # old rows are copied, not recomputed
        ek, ev = rel_l2(kv[:, :d], wk[0]), rel_l2(kv[:, d:], wv[0])
        print(f'layer {l} cached K / V rel-L2 vs fp32 oracle: {ek:.2e} / {ev:.2e}')
        assert ek <= LOGIT_TOL and ev <= LOGIT_TOL
    err = rel_l2(torch.cat([got[t].flatten().float().cpu() for t in cfg.tasks]), torch.cat([want[t].flatten() for t in cfg.tasks]))
    print(f'extend[{schedule}, pyramid={pyramid}] logits rel-L2 vs fp32 oracle: {err:.3e}')
    assert err <= LOGIT_TOL                                                          # north_star: logits rel err <= 1e-2
    if not pyramid:     # size-independent property: append == fresh build on the longer sequence
        with torch.no_grad():
            fresh = _cache_fp(model.build_kv_cache(to_cuda(seq)))
        model.build_kv_cache(to_cuda(head))
        model.extend_kv_cache(new.cuda())
        for a, b in zip(_cache_fp(model.kv_cache), fresh):
            assert a.shape == b.shape and rel_l2(a, b) <= LOGIT_TOL


def test_extend_kv_cache_argument_checks():
    ocfg, cfg = make_configs(num_layers=2, num_ns_tokens=4, schedule='linear_to_ns')
    model = R.OneTransModel(cfg).cuda()
    ev = torch.randn(1, 3, 64, device='cuda')
    with pytest.raises(RuntimeError):
        model.extend_kv_cache(ev)                                                   # no cache yet
    _, seq, _ = O.synthetic_batch(ocfg, 1, (6, 5, 4), seed=1)
    model.build_kv_cache(to_cuda({k: v for k, v in seq.items() if k != 'purchase_seq'}))
    with pytest.raises(ValueError):
        model.extend_kv_cache(ev)                                                   # built without the last sequence
    model.build_kv_cache(to_cuda(seq))
    with pytest.raises(ValueError):
        model.extend_kv_cache(torch.randn(2, 3, 64, device='cuda'))                 # one user only
    sizes = [kv.shape[0] for kv in model.kv_cache['layers'] if kv is not None]
    model.extend_kv_cache(ev[:, :0])                                                # nothing new: unchanged
    assert sizes == [kv.shape[0] for kv in model.kv_cache['layers'] if kv is not None] and model.kv_cache['appended'] == 0
    model.extend_kv_cache(ev)
    assert model.kv_cache['layers'][0].shape[0] == sizes[0] + 3 and model.kv_cache['appended'] == 3
