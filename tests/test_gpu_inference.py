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
