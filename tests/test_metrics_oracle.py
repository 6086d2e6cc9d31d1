"""The metrics oracle against the known-answer examples of the Keras API documentation (tf.keras.metrics.AUC / BinaryAccuracy /
Precision / Recall / BinaryCrossentropy docstrings, TF 2.12) and against scikit-learn's exact ROC-AUC."""
import numpy as np
import pytest

from oracle import metrics_oracle as M


def test_keras_docstring_known_answers():
    # AUC(num_thresholds=3): thresholds [-1e-7, 0.5, 1+1e-7]; tp=[2,1,0] fp=[2,0,0] fn=[0,1,2] tn=[0,2,2] -> 0.75
    y, p = [0, 0, 1, 1], [0, 0.5, 0.3, 0.9]
    assert M.keras_auc(y, p, 3) == pytest.approx(0.75, abs=1e-7)
    assert M.keras_auc_by_thresholds(y, p, 3) == pytest.approx(0.75, abs=1e-7)
    pos, neg = M.keras_auc_state(y, p, 3)
    assert np.cumsum(pos[::-1])[::-1].tolist() == [2, 1, 0] and np.cumsum(neg[::-1])[::-1].tolist() == [2, 0, 0]
    assert M.binary_accuracy([1, 1, 0, 0], [0.98, 1, 0, 0.6]) == pytest.approx(0.75)
    assert M.precision([0, 1, 1, 1], [1, 0, 1, 1]) == pytest.approx(2 / 3, abs=1e-7)
    assert M.recall([0, 1, 1, 1], [1, 0, 1, 1]) == pytest.approx(2 / 3, abs=1e-7)
    assert M.binary_crossentropy([[0, 1], [0, 0]], [[0.6, 0.4], [0.4, 0.6]]) == pytest.approx(0.81492424, abs=1e-6)
    assert M.f1([0, 1, 1, 1], [1, 0, 1, 1]) == pytest.approx(2 / 3, abs=1e-6)
    assert M.precision([0, 0], [0.1, 0.2]) == 0.0                    # div_no_nan


def test_bucketed_update_is_the_threshold_definition():
    rng = np.random.default_rng(0)
    for n in (3, 17, 129, 200):
        p = rng.random(5000).astype(np.float32)
        if (n - 1) & (n - 2) == 0:     # on-threshold predictions: only where k / (n - 1) and p * (n - 1) are exact in fp32 (for
            p[:50] = rng.integers(0, n, 50) / np.float32(n - 1)      # n = 200 Keras' own two update paths round differently there)
        p[50:54] = [0.0, 1.0, 1.5, -0.25]                            # clipped into [0, 1] by the bucketed update
        y = (rng.random(5000) < 0.3).astype(np.float32)
        a, b = M.keras_auc(y, p, n), M.keras_auc_by_thresholds(y, np.clip(p, 0, 1), n)
        assert a == pytest.approx(b, abs=2e-6), (n, a, b)


def test_exact_auc_is_sklearn_roc_auc():
    from sklearn.metrics import roc_auc_score
    rng = np.random.default_rng(1)
    for n, levels in ((1000, None), (5000, 7), (300, 2)):
        p = rng.random(n).astype(np.float32)
        if levels:
            p = (np.floor(p * levels) / levels).astype(np.float32)  # heavy ties
        y = (rng.random(n) < 0.2 + 0.5 * p).astype(np.float32)
        assert M.exact_auc(y, p) == pytest.approx(roc_auc_score(y, p), abs=1e-12)
    assert M.exact_auc([0, 0, 1, 1], [0.1, 0.4, 0.35, 0.8]) == pytest.approx(0.75)     # the scikit-learn docs example
    assert np.isnan(M.exact_auc([1, 1], [0.3, 0.4]))
    # the 200-threshold approximation stays close to the exact value on smooth scores
    p = rng.random(200000).astype(np.float32)
    y = (rng.random(200000) < p).astype(np.float32)
    assert abs(M.keras_auc(y, p) - M.exact_auc(y, p)) < 1e-4


def test_user_auc_weights_by_impressions():
    y = np.array([0, 1, 0, 1, 1, 1, 0, 1, 0, 0], np.float32)
    p = np.array([.1, .9, .8, .2, .5, .6, .3, .3, .2, .1], np.float32)
    u = np.array([0, 0, 1, 1, 2, 2, 3, 3, 3, 3])
    # user 0: 1.0 (2 rows); user 1: 0.0 (2 rows); user 2: one class only, skipped; user 3: pos 0.3 vs neg .3 .2 .1 -> (0.5+1+1)/3 (4 rows)
    want = (1.0 * 2 + 0.0 * 2 + (2.5 / 3) * 4) / 8
    assert M.user_auc(y, p, u, 4) == pytest.approx(want, abs=1e-12)


def test_oracle_reproduces_the_committed_golden_vectors():
    """tests/golden/metrics_golden.json (script: tests/golden/make_metrics_golden.py) - a regression pin of the restatement."""
    import json, os
    G = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'metrics_golden.json')))
    assert set(G) == {'smooth', 'ties', 'edges'}
    for name, g in G.items():
        p, y, u = np.array(g['p'], np.float32), np.array(g['y'], np.float32), np.array(g['user'])
        assert [float(v) for v in p] == g['p']                                         # float32 values survive the JSON round trip
        for nt in (200, 17):
            pos, neg = M.keras_auc_state(y, p, nt)
            assert pos.tolist() == g[f'pos_hist_{nt}'] and neg.tolist() == g[f'neg_hist_{nt}']
            assert M.keras_auc_result(pos, neg) == g[f'keras_auc_{nt}']
        assert M.confusion_counts(y, p) == g['confusion']
        assert M.exact_auc(y, p) == g['exact_auc'] == g['exact_auc_u2'] / g['exact_auc_den']
        assert M.user_auc(y, p, u, g['n_users']) == pytest.approx(g['user_auc'], abs=1e-15)
        assert M.binary_crossentropy(y, p) == pytest.approx(g['logloss'], rel=1e-12)
    e = G['edges']                                                                     # hand-checked entries of the edge case (NT = 17)
    p = np.array(e['p'], np.float32)
    b = np.maximum(np.ceil(p[:12] * np.float32(16)) - 1, 0).astype(int).tolist()
    assert b == [0, 15, 7, 8, 0, 1, 0, 15, 14, 0, 3, 11]
