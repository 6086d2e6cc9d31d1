"""Offline evaluation of a OneTrans model with the surface of the reference's ``OneTransEvaluator`` (OT/evaluate.py:22-361) and
its helpers ``load_model_for_evaluation`` (:364-388) and ``evaluate_model`` (:391-416) - SURVEY.md §8f rank 4.

Same method names, arguments and result keys.  Differences, all forced by the reference not running as written:
* the model is called with two arguments (D7); ``benchmark_performance`` passes both feature dicts (the reference passes
  ``batch[0]`` alone, :187, 206);
* batch times are CUDA-event times on the launching stream (``time.time()`` around an asynchronous launch measures nothing);
* metrics are the streaming kernels of ``recommend_b200.metrics``; ``F1Score`` is 2PR/(P+R) (absent from TF 2.12, D11); next to the
  Keras 200-threshold ``{task}_auc`` the results carry ``{task}_auc_exact`` (tie-aware ROC-AUC over the whole pass);
* ``analyze_feature_importance`` returns the reference's normalised placeholder scores by default (:264-279) and, with
  ``method='ablation'``, the measured drop of the exact ctr AUC when a feature is zeroed;
* the report is the JSON file only: the plots need matplotlib / seaborn, which this image lacks."""
from __future__ import annotations

import json
import time
from pathlib import Path
from typing import Any, Dict, List, Optional, Tuple

import torch

from .config import OneTransConfig
from .metrics import BinaryTaskMetrics, exact_auc
from .model import OneTransModel


def _percentile(xs: List[float], q: float) -> float:
    """numpy's default (linear) percentile of a small host list."""
    s = sorted(xs)
    pos = (len(s) - 1) * q / 100.0
    lo = int(pos)
    hi = min(lo + 1, len(s) - 1)
    return s[lo] + (s[hi] - s[lo]) * (pos - lo)


class OneTransEvaluator:
    def __init__(self, model: OneTransModel, config: OneTransConfig, device: str = 'cuda'):
        self.device = torch.device(device)
        self.model = model.to(self.device).eval()
        self.config = config
        self.metrics = self._create_metrics()                                                   # OT/evaluate.py:30
        self.performance_stats = {'inference_time': [], 'memory_usage': [], 'throughput': []}   # :33-37

    def _create_metrics(self) -> BinaryTaskMetrics:
        return BinaryTaskMetrics(self.config.tasks, self.device)

    def _dataset(self, data_loader, dataset_type: str):
        if dataset_type == 'test':                                                              # :63-68
            return data_loader.get_test_dataset()
        if dataset_type == 'val':
            return data_loader.get_val_dataset()
        return data_loader.get_train_dataset()

    def _to_device(self, batch):
        return tuple({k: v.to(self.device, non_blocking=True) for k, v in part.items()} for part in batch)

    @torch.no_grad()
    def _timed_forward(self, non_seq, seq) -> Tuple[Dict[str, torch.Tensor], torch.cuda.Event, torch.cuda.Event]:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        preds = self.model(non_seq, seq, training=False)
        e1.record()
        return preds, e0, e1

    @torch.no_grad()
    def evaluate_offline(self, data_loader, dataset_type: str = 'test', ablate: Optional[str] = None, rank: int = 0,
                         world_size: int = 1) -> Dict[str, Any]:
        """OT/evaluate.py:58-129: one pass over the dataset; every ``{task}_{auc,accuracy,precision,recall,f1,logloss}`` plus the
        ``'performance'`` block (``total_samples``, ``avg_inference_time_per_batch``, ``throughput_samples_per_second``,
        ``avg_inference_time_per_sample``).  ``world_size > 1`` (one process per GPU, ``torch.distributed`` initialised): rank r takes
        batches ``r, r + world_size, ...`` and the metric states - counts - are summed over the ranks, so every rank returns the
        metrics of the whole dataset; ``{task}_auc_exact`` and the performance block then describe the rank's own shard."""
        self.metrics.reset_states()                                                             # :71-73
        total_samples, events = 0, []
        kept = {t: ([], []) for t in self.config.tasks}
        for i_batch, batch in enumerate(self._dataset(data_loader, dataset_type)):
            if i_batch % world_size != rank:
                continue
            non_seq, seq, labels = self._to_device(batch)
            if ablate is not None:
                for d in (non_seq, seq):
                    if ablate in d:
                        d[ablate] = torch.zeros_like(d[ablate])
            total_samples += next(iter(labels.values())).shape[0]
            preds, e0, e1 = self._timed_forward(non_seq, seq)
            events.append((e0, e1))
            have = [t for t in self.config.tasks if t in preds and t in labels]                 # :91-93
            if len(have) != len(self.config.tasks):
                raise KeyError(f'evaluate_offline: batch lacks labels or predictions for {set(self.config.tasks) - set(have)}')
            self.metrics.update_state(labels, preds)
            for t in have:
                kept[t][0].append(labels[t].reshape(-1).float())
                kept[t][1].append(preds[t].reshape(-1).float())
        if not events:
            raise ValueError('evaluate_offline: empty dataset')
        if world_size > 1:
            self.metrics.all_reduce()
        results: Dict[str, Any] = dict(self.metrics.result())                                   # :108-111
        for t in self.config.tasks:
            results[f'{t}_auc_exact'] = exact_auc(torch.cat(kept[t][0]), torch.cat(kept[t][1]))
        torch.cuda.synchronize(self.device)
        times = [e0.elapsed_time(e1) / 1000.0 for e0, e1 in events]
        avg = sum(times) / len(times)
        results['performance'] = {                                                              # :113-122
            'total_samples': total_samples,
            'avg_inference_time_per_batch': avg,
            'throughput_samples_per_second': total_samples / sum(times),
            'avg_inference_time_per_sample': avg / (total_samples / len(times)),
        }
        return results

    def evaluate_ab_test(self, control_group, treatment_group, metric_name: str = 'ctr_auc') -> Dict[str, Any]:
        """OT/evaluate.py:131-169 (the "significance" is the reference's 1 % rule of thumb, :152)."""
        control_metric = self.evaluate_offline(control_group, 'test').get(metric_name, 0)
        treatment_metric = self.evaluate_offline(treatment_group, 'test').get(metric_name, 0)
        improvement = treatment_metric - control_metric
        pct = (improvement / control_metric) * 100 if control_metric != 0 else 0
        return {'control_metric': control_metric, 'treatment_metric': treatment_metric, 'absolute_improvement': improvement,
                'relative_improvement_percentage': pct, 'is_statistically_significant': abs(pct) > 1.0, 'metric_name': metric_name}

    @torch.no_grad()
    def benchmark_performance(self, data_loader, num_batches: int = 100, warmup_batches: int = 10) -> Dict[str, float]:
        """OT/evaluate.py:171-229: warm-up batches, then per-batch forward time and allocator growth over ``num_batches`` batches."""
        dataset = data_loader.get_test_dataset()
        for i, batch in enumerate(dataset):                                                     # :180-184
            if i >= warmup_batches:
                break
            non_seq, seq, _ = self._to_device(batch)
            self.model(non_seq, seq, training=False)
        events, memory_usages = [], []
        for i, batch in enumerate(dataset):                                                     # :190-213
            if i >= num_batches:
                break
            non_seq, seq, _ = self._to_device(batch)
            before = torch.cuda.memory_allocated(self.device)
            _, e0, e1 = self._timed_forward(non_seq, seq)
            events.append((e0, e1))
            memory_usages.append(torch.cuda.memory_allocated(self.device) - before)
        if not events:
            raise ValueError('benchmark_performance: empty dataset')
        torch.cuda.synchronize(self.device)
        t = [e0.elapsed_time(e1) / 1000.0 for e0, e1 in events]
        mean = sum(t) / len(t)
        std = (sum((x - mean) ** 2 for x in t) / len(t)) ** 0.5
        self.performance_stats['inference_time'].extend(t)
        return {'avg_inference_time_ms': mean * 1000, 'std_inference_time_ms': std * 1000,                       # :216-225
                'p95_inference_time_ms': _percentile(t, 95) * 1000, 'p99_inference_time_ms': _percentile(t, 99) * 1000,
                'throughput_batches_per_second': 1.0 / mean,
                'avg_memory_usage_mb': sum(memory_usages) / len(memory_usages) / (1024 * 1024),
                'max_memory_usage_mb': max(memory_usages) / (1024 * 1024), 'total_batches_tested': num_batches}

    def analyze_feature_importance(self, data_loader, num_samples: int = 1000, method: str = 'reference') -> Dict[str, float]:
        """OT/evaluate.py:231-282.  ``'reference'``: 0.1 per scalar feature, 0.2 per sequence, normalised (:264-279);
        ``'ablation'``: drop of the exact ctr AUC when the feature is zeroed, clipped at 0 and normalised the same way."""
        fc = self.config.feature_config
        scalars = fc['user_features'] + fc['item_features'] + fc['context_features']
        if method == 'reference':
            imp = {**{f: 0.1 for f in scalars}, **{s: 0.2 for s in fc['sequence_features']}}
        elif method == 'ablation':
            key = f'{self.config.tasks[0]}_auc_exact'
            base = self.evaluate_offline(data_loader, 'test')[key]
            imp = {f: max(base - self.evaluate_offline(data_loader, 'test', ablate=f)[key], 0.0) for f in scalars + fc['sequence_features']}
        else:
            raise ValueError(f'unknown method {method!r}')
        total = sum(imp.values())
        return {k: v / total for k, v in imp.items()} if total > 0 else imp

    def generate_evaluation_report(self, data_loader, output_dir: str = './evaluation_reports') -> str:
        """OT/evaluate.py:284-316 without the plots (:318-361 need matplotlib)."""
        out = Path(output_dir)
        out.mkdir(parents=True, exist_ok=True)
        report = {'model_config': self.config.to_dict(),
                  'offline_evaluation': self.evaluate_offline(data_loader),
                  'performance_benchmark': self.benchmark_performance(data_loader),
                  'feature_importance': self.analyze_feature_importance(data_loader),
                  'evaluation_timestamp': time.strftime('%Y-%m-%d %H:%M:%S'),
                  'data_info': data_loader.get_data_info()}
        path = out / 'evaluation_report.json'
        with open(path, 'w') as f:
            json.dump(report, f, indent=2)
        return str(path)


def load_model_for_evaluation(model_path: str, device: str = 'cuda') -> Tuple[OneTransModel, OneTransConfig]:
    """OT/evaluate.py:364-388: ``config.json`` + weights from a directory written by ``OneTransTrainer.save_model``."""
    from . import state
    model_path = Path(model_path)
    if not (model_path / 'config.json').exists():
        raise FileNotFoundError(f'config file not found: {model_path / "config.json"}')
    with open(model_path / 'config.json') as f:
        config = OneTransConfig.from_dict(json.load(f))
    if not (model_path / state.WEIGHTS_FILE).exists():
        raise FileNotFoundError(f'model weights not found: {model_path / state.WEIGHTS_FILE}')
    model = OneTransModel(config)
    state.load_weights(model, model_path / state.WEIGHTS_FILE)
    return model.to(device), config


def evaluate_model(model_path: str, data_loader, output_dir: str = './evaluation_reports') -> Dict:
    """OT/evaluate.py:391-416."""
    model, config = load_model_for_evaluation(model_path)
    evaluator = OneTransEvaluator(model, config)
    return {'model_path': model_path, 'report_path': evaluator.generate_evaluation_report(data_loader, output_dir), 'evaluator': evaluator}


def _cli(argv=None) -> int:
    """``python -m recommend_b200.evaluate`` - the command line of OT/evaluate.py:419-465 (``--model_path``, ``--data_dir``,
    ``--output_dir``, ``--eval_type full|offline|performance|ab_test``); data is generated sample data as in the reference (:445-448)."""
    import argparse
    ap = argparse.ArgumentParser(description='Evaluate a OneTrans model on the sm_100a path')
    ap.add_argument('--model_path', type=str, required=True)
    ap.add_argument('--data_dir', type=str, default=None)
    ap.add_argument('--output_dir', type=str, default='./evaluation_reports')
    ap.add_argument('--eval_type', type=str, default='full', choices=['full', 'offline', 'performance', 'ab_test'])
    args = ap.parse_args(argv)
    if not torch.cuda.is_available():
        print('recommend_b200.evaluate needs a CUDA device (sm_100a); there is no CPU fallback')
        return 2
    from .data import DataLoader
    model, config = load_model_for_evaluation(args.model_path)
    data_loader = DataLoader(config)
    data_loader.train_dataset = data_loader.create_sample_data(seed=1)
    data_loader.test_dataset = data_loader.create_sample_data(seed=3)
    evaluator = OneTransEvaluator(model, config)
    if args.eval_type == 'full':
        print('report:', evaluator.generate_evaluation_report(data_loader, args.output_dir))
    elif args.eval_type == 'offline':
        print(json.dumps(evaluator.evaluate_offline(data_loader), indent=2))
    elif args.eval_type == 'performance':
        print(json.dumps(evaluator.benchmark_performance(data_loader), indent=2))
    else:
        print('ab_test needs a control and a treatment dataset (OT/evaluate.py:464-465); call OneTransEvaluator.evaluate_ab_test')
    return 0


if __name__ == '__main__':
    raise SystemExit(_cli())
