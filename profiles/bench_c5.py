"""BASELINE config 5: inference with a cross-candidate cache of the sequence-side K/V (north_star item 5, PAPER:144-151):
one user x 8192 candidates, NS-token-only queries per candidate.  Prints one JSON line: candidates/s of
(a) stage 2 alone (cache resident), (b) stage 1 + stage 2, (c) the uncached forward over 8192 full sequences.
usage: python profiles/bench_c5.py [C] [steps]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import recommend_b200 as R
from recommend_b200 import _lib
from recommend_b200.data import create_sample_batch

C = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
K = int(sys.argv[2]) if len(sys.argv) > 2 else 10
cfg = R.get_model_config('small'); cfg.num_ns_tokens = 32; cfg.pyramid_schedule = 'linear_to_ns'; cfg.dropout_rate = 0.0
torch.manual_seed(0)
model = R.OneTransModel(cfg).cuda().eval()
ns, sq, _ = create_sample_batch(cfg, C, (170, 170, 170), seed=5)
ns = {k: v.cuda() for k, v in ns.items()}
user_seq = {k: v[:1].cuda().bfloat16() for k, v in sq.items()}                 # ONE user's behaviour sequences


def timed(fn, n):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out


with torch.no_grad():
    model.build_kv_cache(user_seq)
    l0 = _lib.launch_count
    ms_stage2, cached = timed(lambda: model.score_candidates(ns), K)
    launches = (_lib.launch_count - l0) // (K + 2)
    ms_both, _ = timed(lambda: (model.build_kv_cache(user_seq), model.score_candidates(ns))[1], K)
    Cu = min(C, 2048)                                                            # uncached: every candidate re-runs the whole sequence
    full_seq = {k: v.expand(Cu, -1, -1).contiguous() for k, v in user_seq.items()}
    ns_u = {k: v[:Cu] for k, v in ns.items()}
    ms_unc, unc = timed(lambda: model(ns_u, full_seq), max(2, K // 3))
    err = max(float((cached[t][:Cu] - unc[t]).abs().max()) for t in cfg.tasks)
if os.environ.get('C5_PROFILE'):
    from recommend_b200 import ops
    prof = ops.KernelProfiler(); ops.set_profiler(prof)
    with torch.no_grad():
        model.score_candidates(ns)
    torch.cuda.synchronize(); ops.set_profiler(None)
    for (n, t), d in sorted(prof.summary().items(), key=lambda kv: -kv[1]['ms']):
        print(f"{n:24s} {t:28s} n={d['launches']:3d} ms={d['ms']:7.3f} TF={d['flops'] / (d['ms'] * 1e-3) / 1e12:7.1f} GB/s={d['bytes'] / (d['ms'] * 1e-3) / 1e9:7.1f}", file=sys.stderr)
print(json.dumps({'metric': 'OneTrans candidates/sec (cached inference, C5)', 'unit': 'candidates/s', 'candidates': C,
                  'stage2_value': C / (ms_stage2 * 1e-3), 'stage2_ms': ms_stage2, 'stage1_plus_2_value': C / (ms_both * 1e-3),
                  'stage1_plus_2_ms': ms_both, 'uncached_value': Cu / (ms_unc * 1e-3), 'uncached_ms_per_%d' % Cu: ms_unc,
                  'speedup_vs_uncached': (C / ms_both) / (Cu / ms_unc), 'gpu_launches_stage2': launches,
                  'max_abs_prob_diff_cached_vs_uncached': err, 'dtype': 'bf16', 'data': 'synthetic'}))
