"""Evaluation metrics of the OneTrans trainer and evaluator on the GPU.

The reference keeps, per binary task, ``tf.keras.metrics.AUC / BinaryAccuracy / Precision / Recall`` (OT/train.py:95-109)
and adds ``F1Score`` and the ``BinaryCrossentropy`` metric in the evaluator (OT/evaluate.py:39-56); it feeds them with
``update_state(labels[task], predictions[task])`` after every batch (OT/train.py:141-150, 178-187; OT/evaluate.py:91-99),
reads ``result()`` once per epoch / evaluation (OT/train.py:248-249; OT/evaluate.py:109-111) and clears them with
``reset_states()`` (OT/train.py:280-283; OT/evaluate.py:72-73).

Here one kernel pass per batch (``ot_metrics_update``: 8 bytes per (sample, task)) feeds all six metrics of all tasks into a
small int64 state on the device; ``result()`` runs ``ot_metrics_result`` and copies ``n_tasks * 8`` doubles to the host.
``BinaryTaskMetrics`` is that state with the reference's naming (``'{task}_auc'`` ...); the classes ``AUC``,
``BinaryAccuracy``, ``Precision``, ``Recall``, ``F1Score``, ``BinaryCrossentropy`` keep the Keras method surface
(``update_state(y_true, y_pred)`` / ``result()`` / ``reset_states()``) for code written against the reference.

``F1Score`` does not exist in the pinned TensorFlow 2.12; it is ``2PR / (P + R)`` at the same 0.5 threshold
as ``Precision`` / ``Recall`` here.  Keras ``AUC()`` is a 200-threshold approximation; ``exact_auc`` /
``grouped_auc`` give the exact tie-aware ROC-AUC (overall, and per user = UAUC) used for the north_star's AUC-delta check.
No CPU path: CPU tensors raise."""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple, Union

import torch

from . import _lib as L
from . import ops

METRIC_NAMES = ('auc', 'accuracy', 'precision', 'recall', 'f1', 'logloss')


def _as_task_rows(x: Union[torch.Tensor, Dict[str, torch.Tensor]], tasks: Sequence[str], what: str) -> torch.Tensor:
    """``{task: [B, 1] | [B]}`` or an already packed ``[n_tasks, B]`` tensor -> contiguous fp32 ``[n_tasks, B]`` on the GPU."""
    if isinstance(x, dict):
        x = torch.stack([x[t].reshape(-1) for t in tasks])
    if not x.is_cuda:
        raise RuntimeError(f'metrics: {what} must be CUDA tensors (no CPU fallback)')
    if x.dim() != 2 or x.shape[0] != len(tasks):
        raise ValueError(f'metrics: {what} must be [n_tasks={len(tasks)}, B], got {tuple(x.shape)}')
    return x.to(torch.float32).contiguous()


class BinaryTaskMetrics:
    """Streaming AUC / accuracy / precision / recall / F1 / log-loss for every task, one device state.

    ``update_state(labels, predictions)`` takes the dicts the model and the loader produce (``{task: [B, 1]}``) or packed
    ``[n_tasks, B]`` tensors (the layout ``ot_heads_fwd`` writes); ``result()`` returns ``{'ctr_auc': ..., 'ctr_accuracy': ...}``
    as Python floats, named as OT/train.py:95-109 / OT/evaluate.py:39-56 name them."""

    def __init__(self, tasks: Sequence[str], device: Union[str, torch.device] = 'cuda', num_thresholds: int = 200,
                 threshold: float = 0.5):
        if not 3 <= num_thresholds <= L.METRICS_MAX_THRESHOLDS:
            raise ValueError(f'num_thresholds must be in [3, {L.METRICS_MAX_THRESHOLDS}]')
        self.tasks = list(tasks)
        self.num_thresholds, self.threshold = int(num_thresholds), float(threshold)
        self.stride = 2 * self.num_thresholds + L.METRICS_TAIL_WORDS
        self.state = torch.zeros(len(self.tasks), self.stride, dtype=torch.int64, device=device)
        self._result = torch.zeros(len(self.tasks), L.METRICS_RESULT_WORDS, dtype=torch.float64, device=device)

    def _params(self) -> L.MetricsParams:
        p = L.MetricsParams()
        p.n_tasks, p.num_thresholds, p.threshold = len(self.tasks), self.num_thresholds, self.threshold
        p.state, p.state_stride, p.result = self.state.data_ptr(), self.stride, self._result.data_ptr()
        return p

    def update_state(self, labels, predictions) -> None:
        y = _as_task_rows(labels, self.tasks, 'labels')
        pr = _as_task_rows(predictions, self.tasks, 'predictions')
        if y.shape != pr.shape:
            raise ValueError(f'metrics: labels {tuple(y.shape)} vs predictions {tuple(pr.shape)}')
        p = self._params()
        p.probs, p.labels, p.ld, p.B = pr.data_ptr(), y.data_ptr(), pr.stride(0), pr.shape[1]
        ops._run('ot_metrics_update', L.load().ot_metrics_update, p, f'T{len(self.tasks)}', 0.0, 8.0 * pr.numel())

    def result_tensor(self) -> torch.Tensor:
        """``[n_tasks, 8]`` float64 on the device: auc accuracy precision recall f1 logloss count rejected."""
        ops._run('ot_metrics_result', L.load().ot_metrics_result, self._params(), f'T{len(self.tasks)}')
        return self._result

    def result(self) -> Dict[str, float]:
        host = self.result_tensor().cpu()
        out: Dict[str, float] = {}
        for t, task in enumerate(self.tasks):
            if host[t, 7] > 0:
                raise ValueError(f'metrics: task {task!r} saw {int(host[t, 7])} samples with a NaN prediction or a label outside {{0, 1}}')
            for k, name in enumerate(METRIC_NAMES):
                out[f'{task}_{name}'] = float(host[t, k])
        return out

    def counts(self) -> Dict[str, Dict[str, int]]:
        """Raw confusion counts per task (``tp fp tn fn count rejected``) - integers, for exact comparisons."""
        tail = self.state[:, 2 * self.num_thresholds:2 * self.num_thresholds + 6].cpu()
        return {task: dict(zip(('tp', 'fp', 'tn', 'fn', 'count', 'rejected'), (int(v) for v in tail[t]))) for t, task in enumerate(self.tasks)}

    def reset_states(self) -> None:
        self.state.zero_()

    reset_state = reset_states      # Keras >= 2.5 spelling

    def all_reduce(self, group=None) -> None:
        """Data-parallel evaluation: every rank streams its shard of the samples, then the states are summed over the ranks
        (``merge_states_``) and every rank's ``result()`` is the metric of the whole dataset - exact, because the state is
        counts (the reference evaluates on one device, OT/evaluate.py:58-129)."""
        merge_states_(self.state, self.num_thresholds, group)


def merge_states_(state: torch.Tensor, num_thresholds: int, group=None) -> torch.Tensor:
    """Sum metric states ``[n_tasks, 2 * NT + 8]`` (int64) over the ranks of ``group``, in place.  Every word is an integer count
    except word ``2 * NT + 6``, the BCE sum, whose int64 slot holds the bits of a double: it is reduced as float64."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return state
    k = 2 * num_thresholds + 6
    bce = state[:, k].clone().view(torch.float64)
    state[:, k] = 0
    dist.all_reduce(state, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(bce, op=dist.ReduceOp.SUM, group=group)
    state[:, k] = bce.view(torch.int64)
    return state


class _SingleMetric:
    """Keras-style single metric over ``y_true`` / ``y_pred`` of any (equal) shape: a one-task ``BinaryTaskMetrics``."""
    _index = 0

    def __init__(self, name: Optional[str] = None, device: Union[str, torch.device] = 'cuda', **kw):
        self.name = name or type(self).__name__.lower()
        self._m = BinaryTaskMetrics(['m'], device=device, **kw)

    def update_state(self, y_true: torch.Tensor, y_pred: torch.Tensor) -> None:
        self._m.update_state(y_true.reshape(1, -1), y_pred.reshape(1, -1))

    def result(self) -> float:
        return self._m.result()[f'm_{METRIC_NAMES[self._index]}']

    def reset_states(self) -> None:
        self._m.reset_states()

    reset_state = reset_states


class AUC(_SingleMetric):
    """``tf.keras.metrics.AUC(num_thresholds=200, curve='ROC', summation_method='interpolation')``."""
    _index = 0

    def __init__(self, num_thresholds: int = 200, name: Optional[str] = None, device='cuda'):
        super().__init__(name, device, num_thresholds=num_thresholds)


class BinaryAccuracy(_SingleMetric):
    _index = 1

    def __init__(self, name: Optional[str] = None, threshold: float = 0.5, device='cuda'):
        super().__init__(name, device, threshold=threshold)


class Precision(_SingleMetric):
    _index = 2

    def __init__(self, thresholds: Optional[float] = None, name: Optional[str] = None, device='cuda'):
        super().__init__(name, device, threshold=0.5 if thresholds is None else thresholds)


class Recall(Precision):
    _index = 3


class F1Score(Precision):
    _index = 4


class BinaryCrossentropy(_SingleMetric):
    _index = 5


def grouped_auc(labels: torch.Tensor, predictions: torch.Tensor, segment_ids: Optional[torch.Tensor] = None,
                n_segments: int = 1) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Exact tie-aware ROC-AUC per segment.  Returns ``(auc [n_segments] float64 (NaN where a segment lacks a class),
    count [n_segments] int64, positives [n_segments] int64)`` on the device.  ``segment_ids``: int32 in ``[0, n_segments)``."""
    pr = predictions.reshape(-1)
    y = labels.reshape(-1)
    if not (pr.is_cuda and y.is_cuda) or (segment_ids is not None and not segment_ids.is_cuda):
        raise RuntimeError('grouped_auc: CUDA tensors only (no CPU fallback)')
    pr, y = pr.to(torch.float32).contiguous(), y.to(torch.float32).contiguous()
    n = pr.numel()
    if y.numel() != n:
        raise ValueError('grouped_auc: labels and predictions differ in size')
    dev = pr.device
    keys = torch.empty(n, dtype=torch.int64, device=dev)
    rejected = torch.zeros(1, dtype=torch.int32, device=dev)
    stats = torch.zeros(3, n_segments, dtype=torch.int64, device=dev)
    p = L.AucParams()
    p.probs, p.labels, p.n, p.n_segments = pr.data_ptr(), y.data_ptr(), n, n_segments
    if segment_ids is not None:
        seg = segment_ids.reshape(-1).to(torch.int32).contiguous()
        if seg.numel() != n:
            raise ValueError('grouped_auc: segment_ids differ in size')
        p.segment_ids = seg.data_ptr()
    p.keys, p.rejected = keys.data_ptr(), rejected.data_ptr()
    p.seg_count, p.seg_pos, p.seg_sum2 = stats[0].data_ptr(), stats[1].data_ptr(), stats[2].data_ptr()
    lib = L.load()
    ops._run('ot_auc_pack_keys', lib.ot_auc_pack_keys, p, '', 0.0, (16.0 + 4.0 * (segment_ids is not None)) * n)
    keys_sorted = torch.sort(keys).values            # the sort itself is framework plumbing (radix sort of int64 keys)
    p.keys = keys_sorted.data_ptr()
    ops._run('ot_auc_ranksum', lib.ot_auc_ranksum, p, '', 0.0, 8.0 * n)
    if int(rejected.item()):
        raise ValueError(f'grouped_auc: {int(rejected.item())} samples with a NaN prediction, a label outside {{0, 1}} or a segment id outside [0, {n_segments})')
    cnt, pos, sum2 = stats[0], stats[1], stats[2]
    start = torch.cumsum(cnt, 0) - cnt
    u2 = sum2 - 2 * pos * start - pos * (pos + 1)            # 2 * U, integers
    neg = cnt - pos
    auc = u2.to(torch.float64) / (2.0 * (pos * neg).to(torch.float64))
    auc = torch.where((pos > 0) & (neg > 0), auc, torch.full_like(auc, float('nan')))
    return auc, cnt, pos


def exact_auc(labels: torch.Tensor, predictions: torch.Tensor) -> float:
    """Exact ROC-AUC of one prediction vector (what ``sklearn.metrics.roc_auc_score`` returns)."""
    return float(grouped_auc(labels, predictions)[0][0])


def user_auc(labels: torch.Tensor, predictions: torch.Tensor, user_index: torch.Tensor, n_users: int) -> float:
    """UAUC: impression-weighted mean of the per-user exact AUC over the users that have both classes (PAPER's offline
    metric next to AUC).  ``user_index``: dense int32 ids in ``[0, n_users)``."""
    auc, cnt, _ = grouped_auc(labels, predictions, user_index, n_users)
    ok = ~torch.isnan(auc)
    if not bool(ok.any()):
        return float('nan')
    w = cnt[ok].to(torch.float64)
    return float((auc[ok] * w).sum() / w.sum())
