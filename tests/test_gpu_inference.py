"""The serving wrapper (recommend_b200.inference, surface of OT/examples/inference_example.py:21-219): batch == single,
two-stage cached ranking == plain batch inference on the same candidates, stats bookkeeping, model directory loading."""
import json

import pytest
import torch

import recommend_b200 as R

pytestmark = pytest.mark.gpu


def _engine(tmp_path=None):
    cfg = R.get_model_config('small')
    cfg.num_layers, cfg.num_ns_tokens, cfg.max_seq_len, cfg.pyramid_schedule = 2, 4, 12, 'linear_to_ns'
    torch.manual_seed(0)
    model = R.OneTransModel(cfg)
    if tmp_path is not None:
        json.dump(cfg.to_dict(), open(tmp_path / 'config.json', 'w'))
        torch.save(model.state_dict(), tmp_path / 'model_weights.pt')
        return R.OneTransInferenceEngine(tmp_path), cfg
    return R.OneTransInferenceEngine(model), cfg


def _sample(cfg, g, n_events=(9, 12, 5)):
    fc = cfg.feature_config
    mk = lambda names: {n: float(torch.randn((), generator=g)) for n in names}
    seqs = {n: torch.randn(L, 64, generator=g) for n, L in zip(fc['sequence_features'], n_events)}
    return mk(fc['user_features']), mk(fc['item_features']), mk(fc['context_features']), seqs


def test_single_batch_and_cached_ranking_agree(tmp_path):
    eng, cfg = _engine(tmp_path)
    g = torch.Generator().manual_seed(3)
    user, _, ctx, seqs = _sample(cfg, g)
    items = [_sample(cfg, g)[1] for _ in range(6)]
    batch = [(user, it, ctx, seqs) for it in items]
    out_b = eng.batch_inference(batch)
    out_s = eng.single_inference(*batch[2])
    assert set(out_b[0]) == set(cfg.tasks) and len(out_b) == 6
    assert all(abs(out_s[t] - out_b[2][t]) < 2e-3 for t in cfg.tasks)       # same rows, different batch size: bf16 tile order only
    cand = {}
    for name in cfg.ns_features:
        src = user if name in user else ctx
        cand[name] = [it[name] if name in it else src[name] for it in items]
    ranked = eng.rank_candidates(seqs, cand)
    for t in cfg.tasks:
        for i in range(6):
            assert abs(ranked[t][i] - out_b[i][t]) < 1e-2, (t, i, ranked[t][i], out_b[i][t])
    st = eng.get_stats()
    assert st['total_requests'] == 6 + 1 + 6 and st['failed_requests'] == 0 and st['success_rate'] == 100.0 and st['avg_latency_ms'] > 0
    eng.reset_stats()
    assert eng.get_stats()['total_requests'] == 0


def test_preprocess_pads_and_truncates_like_the_reference():
    eng, cfg = _engine()
    g = torch.Generator().manual_seed(1)
    user, item, ctx, _ = _sample(cfg, g)
    seqs = {'click_seq': torch.ones(5, 64), 'cart_seq': torch.ones(20, 64)}
    non_seq, seq = eng.preprocess_input(user, item, ctx, seqs)
    assert set(non_seq) == set(user) | set(item) | set(ctx)
    assert seq['click_seq'].shape == (12, 64) and seq['click_seq'][:7].abs().sum() == 0 and seq['click_seq'][7:].eq(1).all()   # left pad
    assert seq['cart_seq'].shape == (12, 64)                                                                                    # last 12 events
    with pytest.raises(FileNotFoundError):
        R.OneTransInferenceEngine('/nonexistent/model/dir')


def test_wrapper_outputs_equal_the_oracle():
    """The serving wrapper against the fp32 oracle (not against itself): ``batch_inference`` and the two-stage ``rank_candidates``
    on the wrapper's own preprocessed (left-padded) inputs, GEMM weights bf16-representable on both sides; logits rel-L2 <= 1e-2."""
    from oracle import onetrans_oracle as O
    from tests.helpers import rel_l2
    eng, cfg = _engine()
    P = R.export_reference_style_params(eng.model)
    P = {k: (v.to(torch.bfloat16).float() if (v.dim() >= 2 and 'ns_tokenizer' not in k and 'task_heads' not in k and 'sep_embedding' not in k) else v.float())
         for k, v in P.items()}
    R.load_reference_style_params(eng.model, P)
    ocfg = O.OracleConfig(hidden_dim=cfg.hidden_dim, num_layers=cfg.num_layers, num_heads=cfg.num_heads, ffn_dim=cfg.ffn_dim,
                          num_ns_tokens=cfg.num_ns_tokens, dropout_rate=0.0)
    L0 = 3 * cfg.max_seq_len + 2 + cfg.num_ns_tokens
    ocfg.pyramid_keep_lens = R.resolve_keep_lens(cfg, L0)
    g = torch.Generator().manual_seed(8)
    n = 48
    batch = [_sample(cfg, g, n_events=(3 + i % 9, 12, 1 + i % 5)) for i in range(n)]
    batch = [(u, it, c, {k: v.to(torch.bfloat16).float() for k, v in s.items()}) for u, it, c, s in batch]
    out = eng.batch_inference(batch)
    pre = [eng.preprocess_input(*b) for b in batch]
    non_seq = {k: torch.tensor([[float(p[0][k])] for p in pre]) for k in cfg.ns_features}
    seq = {k: torch.stack([p[1][k].float().cpu() for p in pre]) for k in cfg.feature_config['sequence_features']}
    want = O.model_forward(P, ocfg, non_seq, seq, return_logits=True)
    lo = torch.cat([want[t].flatten() for t in cfg.tasks]).double()
    lg = torch.cat([torch.logit(torch.tensor([o[t] for o in out], dtype=torch.float64)) for t in cfg.tasks])
    e = rel_l2(lg, lo)
    print(f'batch_inference vs fp32 oracle: logits rel-L2 {e:.3e}')
    assert e <= 1e-2
    # two-stage ranking of one user's candidates against the oracle's full forward on the same rows
    user, _, ctx, seqs = batch[0]
    fixed = {**user, **ctx}                      # the user and the request context are shared, the item features vary per candidate
    cand = {name: [float(fixed[name]) if name in fixed else float(b[1][name]) for b in batch] for name in cfg.ns_features}
    ranked = eng.rank_candidates(seqs, cand)
    _, seq1 = eng.preprocess_input(user, batch[0][1], ctx, seqs)
    seq_c = {k: seq1[k].float().cpu().unsqueeze(0).expand(n, -1, -1).contiguous() for k in cfg.feature_config['sequence_features']}
    ns_c = {k: torch.tensor(cand[k], dtype=torch.float32).reshape(n, 1) for k in cfg.ns_features}
    want = O.model_forward(P, ocfg, ns_c, seq_c, return_logits=True)
    lo = torch.cat([want[t].flatten() for t in cfg.tasks]).double()
    lg = torch.cat([torch.logit(torch.tensor(ranked[t], dtype=torch.float64)) for t in cfg.tasks])
    e = rel_l2(lg, lo)
    print(f'rank_candidates vs fp32 oracle: logits rel-L2 {e:.3e}')
    assert e <= 1e-2
