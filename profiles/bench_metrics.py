"""Evaluation-metric kernels against the HBM roofline: ``ot_metrics_update`` streams 8 B per (sample, task),
``ot_auc_pack_keys`` 16 B per sample, ``ot_auc_ranksum`` 8 B per sample.  Inputs are larger than L2 (126 MB).
usage: python profiles/bench_metrics.py [log2_samples_per_task] [steps]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recommend_b200 import metrics as GM, ops

LOG2 = int(sys.argv[1]) if len(sys.argv) > 1 else 26
K = int(sys.argv[2]) if len(sys.argv) > 2 else 10
B, T = 1 << LOG2, 2
peaks = {}
try:
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))
except Exception:
    pass
hbm_peak = float(peaks.get('hbm_gbs', 6534.8))
g = torch.Generator(device='cuda').manual_seed(0)
probs = torch.rand(T, B, device='cuda', generator=g)
labels = (torch.rand(T, B, device='cuda', generator=g) < probs).float()


def timed(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


m = GM.BinaryTaskMetrics(['ctr', 'cvr'])
ms_update = timed(lambda: m.update_state(labels, probs), K)
m.reset_states(); m.update_state(labels, probs)
res = m.result()
n = 1 << min(LOG2, 25)
prof = ops.KernelProfiler(); ops.set_profiler(prof)
for _ in range(4):
    auc = GM.exact_auc(labels[0, :n], probs[0, :n])
torch.cuda.synchronize(); ops.set_profiler(None)
per = {k[0]: v['ms'] / v['launches'] for k, v in prof.summary().items()}
ms_total = timed(lambda: GM.grouped_auc(labels[0, :n], probs[0, :n]), 3)
gbps = 8.0 * T * B / (ms_update * 1e-3) / 1e9
print(json.dumps({'metric': 'streaming metrics update (AUC/accuracy/precision/recall/F1/logloss, 2 tasks)', 'samples_per_task': B,
                  'ms_per_update': ms_update, 'value': T * B / (ms_update * 1e-3), 'unit': 'sample-tasks/s',
                  'roofline': {'bound': 'hbm', 'achieved': gbps, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': gbps / hbm_peak, 'traffic': None},
                  'ctr_auc_keras200': res['ctr_auc'], 'exact_auc': {'samples': n, 'value': auc, 'ms_pack': per.get('ot_auc_pack_keys'),
                  'ms_ranksum': per.get('ot_auc_ranksum'), 'ms_total_with_sort': ms_total,
                  'pack_gbps': 16.0 * n / (per.get('ot_auc_pack_keys', 1) * 1e-3) / 1e9, 'ranksum_gbps': 8.0 * n / (per.get('ot_auc_ranksum', 1) * 1e-3) / 1e9},
                  'dtype': 'int64 counts / fp32 inputs', 'data': 'synthetic'}))
