"""CPU oracle for the evaluation metrics — TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as onetrans_oracle.py: only
``tests/`` may import it, as the checker).

numpy restatement of the Keras 2.12 metrics the reference instantiates (OT/train.py:95-109, OT/evaluate.py:39-56).  The
arithmetic lives in the third-party dependency ``tensorflow==2.12.0`` (OT/requirements.txt:1; not under /root/reference, not
installable here), so the published algorithm is restated and pinned by the known-answer examples of the Keras API
documentation for each class (tests/test_metrics_oracle.py): ``AUC(num_thresholds=3)`` on ``[0,0,1,1] / [0,0.5,0.3,0.9]`` = 0.75,
``BinaryAccuracy`` on ``[1,1,0,0] / [0.98,1,0,0.6]`` = 0.75, ``Precision`` / ``Recall`` on ``[0,1,1,1] / [1,0,1,1]`` = 2/3,
``BinaryCrossentropy`` on ``[[0,1],[0,0]] / [[0.6,0.4],[0.4,0.6]]`` = 0.81492424; the exact AUC is pinned by
``sklearn.metrics.roc_auc_score`` (an independent implementation present in the image).

Keras rules followed
  * ``AUC.__init__``: thresholds ``[(i + 1) / (n - 1) for i in range(n - 2)]`` framed by ``0 - eps`` and ``1 + eps``, eps = 1e-7;
    they are evenly spaced, so ``update_state`` takes the bucketed path of ``metrics_utils.update_confusion_matrix_variables``:
    ``bucket = relu(ceil(clip(p, 0, 1) * (n - 1)) - 1)`` in fp32, ``unsorted_segment_sum`` of the labels / (1 - labels) per
    bucket, reverse ``cumsum`` -> tp / fp per threshold, ``fn = P - tp``, ``tn = N - fp``  (``keras_auc_state``)
  * ``AUC.result`` (ROC, interpolation): ``recall = div_no_nan(tp, tp + fn)``, ``fpr = div_no_nan(fp, fp + tn)``,
    ``sum((fpr[:-1] - fpr[1:]) * (recall[:-1] + recall[1:]) / 2)`` in fp32  (``keras_auc_result``)
  * ``keras_auc_by_thresholds`` is the definition the bucketed path implements (``p > t_i`` against the threshold list)
  * ``BinaryAccuracy``: ``mean(y == cast(p > 0.5))``;  ``Precision`` / ``Recall``: counts at ``p > 0.5``, ``div_no_nan``
  * ``BinaryCrossentropy`` metric: mean over samples of ``backend.binary_crossentropy`` (clip to ``[eps, 1 - eps]``, ``log(. + eps)``)
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

EPS = 1e-7


def _div_no_nan(a, b):
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    return np.where(b == 0, np.float32(0), a / np.where(b == 0, np.float32(1), b)).astype(np.float32)


def keras_auc_state(y_true: np.ndarray, y_pred: np.ndarray, num_thresholds: int = 200) -> Tuple[np.ndarray, np.ndarray]:
    """Per-bucket positive / negative counts (int64 ``[num_thresholds]`` each) of the bucketed update."""
    p = np.clip(np.asarray(y_pred, np.float32).reshape(-1), np.float32(0), np.float32(1))
    y = np.asarray(y_true, np.float32).reshape(-1)
    b = np.ceil(p * np.float32(num_thresholds - 1)).astype(np.float32) - np.float32(1)
    b = np.maximum(b, 0).astype(np.int64)
    pos = np.bincount(b[y != 0], minlength=num_thresholds).astype(np.int64)
    neg = np.bincount(b[y == 0], minlength=num_thresholds).astype(np.int64)
    return pos, neg


def keras_auc_result(pos: np.ndarray, neg: np.ndarray) -> float:
    tp = np.cumsum(pos[::-1])[::-1].astype(np.float32)
    fp = np.cumsum(neg[::-1])[::-1].astype(np.float32)
    P, N = np.float32(pos.sum()), np.float32(neg.sum())
    fn, tn = P - tp, N - fp
    recall = _div_no_nan(tp, tp + fn)
    fpr = _div_no_nan(fp, fp + tn)
    return float(np.sum((fpr[:-1] - fpr[1:]) * ((recall[:-1] + recall[1:]) / np.float32(2)), dtype=np.float32))


def keras_auc(y_true, y_pred, num_thresholds: int = 200) -> float:
    return keras_auc_result(*keras_auc_state(y_true, y_pred, num_thresholds))


def keras_auc_by_thresholds(y_true, y_pred, num_thresholds: int = 200) -> float:
    """The same metric from its definition: confusion counts of ``p > t`` for every threshold of the list."""
    n = num_thresholds
    th = np.array([0.0 - EPS] + [(i + 1) / (n - 1) for i in range(n - 2)] + [1.0 + EPS], np.float32)
    p = np.asarray(y_pred, np.float32).reshape(1, -1)
    y = np.asarray(y_true, np.float32).reshape(1, -1) != 0
    above = p > th.reshape(-1, 1)
    tp = (above & y).sum(1).astype(np.float32)
    fp = (above & ~y).sum(1).astype(np.float32)
    fn = (~above & y).sum(1).astype(np.float32)
    tn = (~above & ~y).sum(1).astype(np.float32)
    recall = _div_no_nan(tp, tp + fn)
    fpr = _div_no_nan(fp, fp + tn)
    return float(np.sum((fpr[:-1] - fpr[1:]) * ((recall[:-1] + recall[1:]) / np.float32(2)), dtype=np.float32))


def confusion_counts(y_true, y_pred, threshold: float = 0.5) -> Dict[str, int]:
    p = np.asarray(y_pred, np.float32).reshape(-1) > np.float32(threshold)
    y = np.asarray(y_true, np.float32).reshape(-1) != 0
    return dict(tp=int((p & y).sum()), fp=int((p & ~y).sum()), tn=int((~p & ~y).sum()), fn=int((~p & y).sum()), count=int(y.size))


def binary_accuracy(y_true, y_pred, threshold: float = 0.5) -> float:
    p = (np.asarray(y_pred, np.float32).reshape(-1) > np.float32(threshold)).astype(np.float32)
    return float(np.mean(np.asarray(y_true, np.float32).reshape(-1) == p))


def precision(y_true, y_pred, threshold: float = 0.5) -> float:
    c = confusion_counts(y_true, y_pred, threshold)
    return float(_div_no_nan(c['tp'], c['tp'] + c['fp']))


def recall(y_true, y_pred, threshold: float = 0.5) -> float:
    c = confusion_counts(y_true, y_pred, threshold)
    return float(_div_no_nan(c['tp'], c['tp'] + c['fn']))


def f1(y_true, y_pred, threshold: float = 0.5) -> float:
    """Not in TF 2.12: ``2PR / (P + R)`` at the Precision / Recall threshold."""
    pr, rc = np.float32(precision(y_true, y_pred, threshold)), np.float32(recall(y_true, y_pred, threshold))
    return float(_div_no_nan(np.float32(2) * pr * rc, pr + rc))


def binary_crossentropy(y_true, y_pred) -> float:
    y = np.asarray(y_true, np.float32).reshape(-1)
    p = np.clip(np.asarray(y_pred, np.float32).reshape(-1), np.float32(EPS), np.float32(1.0 - EPS))
    bce = y * np.log(p + np.float32(EPS)) + (np.float32(1) - y) * np.log(np.float32(1) - p + np.float32(EPS))
    return float(np.mean(-bce.astype(np.float64)))


def exact_auc(y_true, y_pred) -> float:
    """Tie-aware ROC-AUC from mid-ranks (Mann-Whitney U / (P N)); NaN when a class is missing."""
    y = np.asarray(y_true).reshape(-1) != 0
    p = np.asarray(y_pred, np.float32).reshape(-1)
    P, N = int(y.sum()), int((~y).sum())
    if P == 0 or N == 0:
        return float('nan')
    order = np.argsort(p, kind='stable')
    ps = p[order]
    first = np.searchsorted(ps, ps, side='left')
    last = np.searchsorted(ps, ps, side='right')
    rank2 = np.empty(p.size, np.int64)                       # twice the 1-based mid-rank
    rank2[order] = first + last + 1
    u2 = int(rank2[y].sum()) - P * (P + 1)
    return u2 / (2.0 * P * N)


def user_auc(y_true, y_pred, user_index, n_users: int) -> float:
    """Impression-weighted mean of the per-user exact AUC over users with both classes."""
    y = np.asarray(y_true).reshape(-1)
    p = np.asarray(y_pred).reshape(-1)
    u = np.asarray(user_index).reshape(-1)
    num = den = 0.0
    for k in range(n_users):
        m = u == k
        a = exact_auc(y[m], p[m]) if m.any() else float('nan')
        if a == a:
            num += a * m.sum()
            den += m.sum()
    return num / den if den else float('nan')
