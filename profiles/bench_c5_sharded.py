"""BASELINE config 5 on N GPUs: one user x C candidates, candidates sharded over the ranks, every rank builds the
user's K/V cache, probabilities all-gathered.  Launch: torchrun --nproc-per-node N profiles/bench_c5_sharded.py [C] [steps].
Rank 0 prints one JSON line: whole-job candidates/s (CUDA events, max over ranks) and the largest difference to the
unsharded scoring of the same candidates on one GPU."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import recommend_b200 as R
from recommend_b200.data import create_sample_batch

C = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
K = int(sys.argv[2]) if len(sys.argv) > 2 else 10
rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
if world > 1:
    dist.init_process_group('nccl')
cfg = R.get_model_config('small'); cfg.num_ns_tokens = 32; cfg.pyramid_schedule = 'linear_to_ns'; cfg.dropout_rate = 0.0
torch.manual_seed(0)                                                           # same weights on every rank
model = R.OneTransModel(cfg).cuda().eval()
ns, sq, _ = create_sample_batch(cfg, C, (170, 170, 170), seed=5)
ns = {k: v.cuda() for k, v in ns.items()}
user_seq = {k: v[:1].cuda().bfloat16() for k, v in sq.items()}
run = lambda: R.score_candidates_sharded(model, user_seq, ns, world, rank)
for _ in range(3):
    out = run()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(K):
    out = run()
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / K], device='cuda')
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
with torch.no_grad():
    model.build_kv_cache(user_seq)
    one = model.score_candidates(ns)
diff = max(float((out[t] - one[t].float()).abs().max()) for t in cfg.tasks)
if rank == 0:
    print(json.dumps({'metric': 'OneTrans candidates/sec (cached inference, C5, candidates sharded)', 'unit': 'candidates/s', 'n_gpus': world,
                      'candidates': C, 'value': C / (float(ms) * 1e-3), 'ms_per_request': float(ms), 'steps': K,
                      'max_abs_prob_diff_vs_one_gpu': diff, 'scaling': 'strong', 'dtype': 'bf16', 'data': 'synthetic'}))
if world > 1:
    dist.destroy_process_group()
