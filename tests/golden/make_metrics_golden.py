"""Generates tests/golden/metrics_golden.json from oracle/metrics_oracle.py (run: python -m tests.golden.make_metrics_golden).

The Keras metric classes live in TensorFlow 2.12, which is neither vendored in the reference nor installable here, so these vectors
pin the oracle's restatement against regressions (the oracle itself is pinned by the Keras documentation's known-answer examples and
by scikit-learn in tests/test_metrics_oracle.py) and give the GPU tests fixed integer targets: bucket histograms, confusion counts
and the exact AUC as a fraction ``u2 / (2 P N)``.  Inputs are stored as the float32 values' exact decimal expansions."""
import json
import os

import numpy as np

from oracle import metrics_oracle as M


def make_inputs(name):
    rng = np.random.default_rng({'smooth': 1, 'ties': 2, 'edges': 3}[name])
    n = 257
    p = rng.random(n).astype(np.float32)
    if name == 'ties':
        p = (np.floor(p * 6) / 6).astype(np.float32)
    if name == 'edges':                                     # clip range, bucket edges of NT = 17, the 0.5 threshold, tiny values
        p[:12] = np.array([0.0, 1.0, 0.5, np.nextafter(np.float32(0.5), np.float32(1)), 0.0625, 0.125, 1e-8, 1.0 - 6e-8, 0.9375, 1e-7, 0.25, 0.75],
                          np.float32)
    y = (rng.random(n) < 0.2 + 0.6 * p).astype(np.float32)
    u = np.minimum((rng.pareto(1.0, n) * 3).astype(np.int64), 9).astype(np.int32)
    return p, y, u


def run_case(name):
    p, y, u = make_inputs(name)
    out = {'p': [float(v) for v in p], 'y': [int(v) for v in y], 'user': [int(v) for v in u], 'n_users': 10}
    for nt in (200, 17):
        pos, neg = M.keras_auc_state(y, p, nt)
        out[f'pos_hist_{nt}'] = pos.tolist()
        out[f'neg_hist_{nt}'] = neg.tolist()
        out[f'keras_auc_{nt}'] = M.keras_auc_result(pos, neg)
    out['confusion'] = M.confusion_counts(y, p)
    out['accuracy'], out['precision'], out['recall'], out['f1'] = M.binary_accuracy(y, p), M.precision(y, p), M.recall(y, p), M.f1(y, p)
    out['logloss'] = M.binary_crossentropy(y, p)
    out['exact_auc'] = M.exact_auc(y, p)
    P, N = int(y.sum()), int((1 - y).sum())
    out['exact_auc_u2'] = int(round(out['exact_auc'] * 2 * P * N))         # an integer by construction (twice the U statistic)
    out['exact_auc_den'] = 2 * P * N
    out['user_auc'] = M.user_auc(y, p, u, 10)
    per = []
    for k in range(10):
        m = u == k
        a = M.exact_auc(y[m], p[m]) if m.any() else float('nan')
        per.append(None if a != a else a)
    out['per_user_auc'] = per
    return out


if __name__ == '__main__':
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'metrics_golden.json')
    with open(path, 'w') as f:
        json.dump({name: run_case(name) for name in ('smooth', 'ties', 'edges')}, f)
    print('wrote', path, os.path.getsize(path), 'bytes')
