"""Streaming metrics kernels (ot_metrics_update / ot_metrics_result / ot_auc_pack_keys / ot_auc_ranksum) against the numpy
oracle: integer state bit-exact, fp32 results to 2e-6, exact AUC equal as rationals (to 1e-12)."""
import numpy as np
import pytest
import torch

from recommend_b200 import metrics as GM
from oracle import metrics_oracle as M

pytestmark = pytest.mark.gpu


def _data(n, seed, levels=None, tasks=2):
    rng = np.random.default_rng(seed)
    p = rng.random((tasks, n)).astype(np.float32)
    if levels:
        p = (np.floor(p * levels) / levels).astype(np.float32)
    y = (rng.random((tasks, n)) < 0.15 + 0.6 * p).astype(np.float32)
    return y, p


@pytest.mark.parametrize('n,nt', [(1, 200), (37, 200), (4096, 200), (300001, 200), (5000, 3), (5000, 512)])
def test_streaming_metrics_match_the_oracle(n, nt):
    y, p = _data(n, n + nt)
    p[0, :4] = [0.0, 1.0, 0.5, 0.5000001][:min(4, n)]
    m = GM.BinaryTaskMetrics(['ctr', 'cvr'], num_thresholds=nt)
    cuts = sorted({0, n // 3, n // 2, n})                       # several update_state calls accumulate into one state
    for a, b in zip(cuts[:-1], cuts[1:]):
        m.update_state({'ctr': torch.from_numpy(y[0, a:b]).cuda().reshape(-1, 1), 'cvr': torch.from_numpy(y[1, a:b]).cuda().reshape(-1, 1)},
                       {'ctr': torch.from_numpy(p[0, a:b]).cuda().reshape(-1, 1), 'cvr': torch.from_numpy(p[1, a:b]).cuda().reshape(-1, 1)})
    res = m.result()
    state = m.state.cpu().numpy()
    for t, task in enumerate(('ctr', 'cvr')):
        pos, neg = M.keras_auc_state(y[t], p[t], nt)
        assert np.array_equal(state[t, :nt], pos) and np.array_equal(state[t, nt:2 * nt], neg)          # bit-exact buckets
        want = M.confusion_counts(y[t], p[t])
        got = m.counts()[task]
        assert {k: got[k] for k in want} == want and got['rejected'] == 0
        assert res[f'{task}_auc'] == pytest.approx(M.keras_auc_result(pos, neg), abs=2e-6)
        assert res[f'{task}_accuracy'] == pytest.approx(M.binary_accuracy(y[t], p[t]), abs=1e-12)
        assert res[f'{task}_precision'] == pytest.approx(M.precision(y[t], p[t]), abs=2e-7)
        assert res[f'{task}_recall'] == pytest.approx(M.recall(y[t], p[t]), abs=2e-7)
        assert res[f'{task}_f1'] == pytest.approx(M.f1(y[t], p[t]), abs=5e-7)
        assert res[f'{task}_logloss'] == pytest.approx(M.binary_crossentropy(y[t], p[t]), rel=2e-6)
    m.reset_states()
    assert int(m.state.abs().sum()) == 0


def test_keras_docstring_examples_on_the_gpu():
    dev = 'cuda'
    a = GM.AUC(num_thresholds=3)
    a.update_state(torch.tensor([0., 0, 1, 1], device=dev), torch.tensor([0, 0.5, 0.3, 0.9], device=dev))
    assert a.result() == pytest.approx(0.75, abs=1e-7)
    acc = GM.BinaryAccuracy()
    acc.update_state(torch.tensor([[1.], [1], [0], [0]], device=dev), torch.tensor([[0.98], [1], [0], [0.6]], device=dev))
    assert acc.result() == pytest.approx(0.75)
    pr, rc, f1 = GM.Precision(), GM.Recall(), GM.F1Score()
    for mt in (pr, rc, f1):
        mt.update_state(torch.tensor([0., 1, 1, 1], device=dev), torch.tensor([1., 0, 1, 1], device=dev))
        assert mt.result() == pytest.approx(2 / 3, abs=1e-6)
    ll = GM.BinaryCrossentropy()
    ll.update_state(torch.tensor([[0., 1], [0, 0]], device=dev), torch.tensor([[0.6, 0.4], [0.4, 0.6]], device=dev))
    assert ll.result() == pytest.approx(0.81492424, abs=1e-6)
    acc.reset_states()
    assert acc.result() == 0.0                                   # empty state: div_no_nan


def test_rejects_nan_and_nonbinary_and_cpu():
    m = GM.BinaryTaskMetrics(['ctr'])
    m.update_state(torch.tensor([[0., 1, 0.5, 1]], device='cuda'), torch.tensor([[0.2, float('nan'), 0.3, 0.9]], device='cuda'))
    assert m.counts()['ctr']['rejected'] == 2 and m.counts()['ctr']['count'] == 2
    with pytest.raises(ValueError):
        m.result()
    with pytest.raises(RuntimeError):
        m.update_state(torch.zeros(1, 4), torch.zeros(1, 4))
    with pytest.raises(ValueError):
        GM.exact_auc(torch.tensor([0., 2.], device='cuda'), torch.tensor([0.1, 0.2], device='cuda'))


@pytest.mark.parametrize('n,levels', [(2, None), (1000, None), (70000, 9), (262144 + 77, 1000), (300, 2)])
def test_exact_auc_matches_midrank_oracle(n, levels):
    y, p = _data(n, 7 * n, levels, tasks=1)
    y[0, 0], y[0, -1] = 0.0, 1.0
    got = GM.exact_auc(torch.from_numpy(y[0]).cuda(), torch.from_numpy(p[0]).cuda())
    assert got == pytest.approx(M.exact_auc(y[0], p[0]), abs=1e-12)


def test_user_auc_matches_the_oracle():
    rng = np.random.default_rng(5)
    n, n_users = 50000, 600
    u = np.minimum((rng.pareto(1.2, n) * 20).astype(np.int64), n_users - 1).astype(np.int32)      # a few heavy users, a long tail, some empty
    u[:300] = 7                                                                                  # a segment longer than a CTA
    y, p = _data(n, 11, levels=50, tasks=1)
    auc, cnt, pos = GM.grouped_auc(torch.from_numpy(y[0]).cuda(), torch.from_numpy(p[0]).cuda(), torch.from_numpy(u).cuda(), n_users)
    auc, cnt, pos = auc.cpu().numpy(), cnt.cpu().numpy(), pos.cpu().numpy()
    for k in range(n_users):
        mk = u == k
        assert cnt[k] == mk.sum() and pos[k] == int(y[0][mk].sum())
        want = M.exact_auc(y[0][mk], p[0][mk]) if mk.any() else float('nan')
        assert (np.isnan(want) and np.isnan(auc[k])) or auc[k] == pytest.approx(want, abs=1e-12), k
    got = GM.user_auc(torch.from_numpy(y[0]).cuda(), torch.from_numpy(p[0]).cuda(), torch.from_numpy(u).cuda(), n_users)
    assert got == pytest.approx(M.user_auc(y[0], p[0], u, n_users), abs=1e-12)
    with pytest.raises(ValueError):
        GM.grouped_auc(torch.from_numpy(y[0]).cuda(), torch.from_numpy(p[0]).cuda(), torch.from_numpy(u).cuda(), 10)   # ids outside the range


def test_kernels_reproduce_the_committed_golden_vectors():
    """tests/golden/metrics_golden.json: integer targets (bucket histograms, confusion counts, twice the U statistic) bit-exact."""
    import json, os
    G = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'metrics_golden.json')))
    for name, g in G.items():
        p = torch.tensor(g['p'], dtype=torch.float32, device='cuda')
        y = torch.tensor(g['y'], dtype=torch.float32, device='cuda')
        u = torch.tensor(g['user'], dtype=torch.int32, device='cuda')
        for nt in (200, 17):
            m = GM.BinaryTaskMetrics(['t'], num_thresholds=nt)
            m.update_state(y.reshape(1, -1), p.reshape(1, -1))
            st = m.state.cpu().numpy()[0]
            assert st[:nt].tolist() == g[f'pos_hist_{nt}'] and st[nt:2 * nt].tolist() == g[f'neg_hist_{nt}'], (name, nt)
            res = m.result()
            assert res['t_auc'] == pytest.approx(g[f'keras_auc_{nt}'], abs=2e-6)
            c = m.counts()['t']
            assert {k: c[k] for k in g['confusion']} == g['confusion']
        assert res['t_accuracy'] == pytest.approx(g['accuracy'], abs=1e-12) and res['t_f1'] == pytest.approx(g['f1'], abs=5e-7)
        assert res['t_logloss'] == pytest.approx(g['logloss'], rel=2e-6)
        auc, cnt, pos = GM.grouped_auc(y, p)
        P, n = int(pos[0]), int(cnt[0])
        assert 2 * P * (n - P) == g['exact_auc_den'] and float(auc[0]) == g['exact_auc']
        assert GM.user_auc(y, p, u, g['n_users']) == pytest.approx(g['user_auc'], abs=1e-12)
        per = GM.grouped_auc(y, p, u, g['n_users'])[0].cpu().tolist()
        for a, b in zip(per, g['per_user_auc']):
            assert (b is None and a != a) or a == pytest.approx(b, abs=1e-12)
