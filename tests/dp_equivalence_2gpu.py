"""SURVEY T10 (second half) on hardware: the gradients of the data-parallel step on 2 GPUs - per-block all-reduces (NCCL, AVG)
issued under the backward pass - equal the gradients of the same GLOBAL batch on one GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/dp_equivalence_2gpu.py

Every rank builds the same model (seed 0; parameters are also broadcast from rank 0), takes its half of one global batch of 2 * B
samples, runs ``train_step(world_size=2, overlap_reduce=True)`` (dropout off: the counter-based mask is a function of the LOCAL row
index, so halves and whole would draw different masks).  Rank 0 then runs the whole global batch alone and compares the flat
gradient buffers.  The summed-BCE loss is a mean over the batch, so mean-of-halves == whole exactly in real arithmetic (the halved
batch only rescales the upstream gradient by a power of two).  What differs is summation ORDER: the weight-gradient kernels flush
fp32 atomics and the attention backward accumulates dQ partials with bf16 vector reductions, so two runs of the SAME one-GPU step
already differ at the 1e-3 .. 1e-2 level per tensor.  The test therefore measures that run-to-run noise (the global batch twice on
one GPU) and requires the data-parallel gradients to sit within 3x of it, and within half the per-tensor tolerance the parity tests
hold the kernels to against the oracle (1.5e-2 of 3e-2)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import recommend_b200 as R
from recommend_b200.data import create_sample_batch
from recommend_b200.train import FlatGradBuffer, broadcast_parameters, train_step


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    cfg = R.get_model_config('small')
    cfg.num_ns_tokens, cfg.pyramid_schedule, cfg.dropout_rate = 32, 'linear_to_ns', 0.0
    torch.manual_seed(0)
    model = R.OneTransModel(cfg).to(dev)
    broadcast_parameters(model, world)
    grads = FlatGradBuffer(model.parameters())
    B = int(os.environ.get('DP_TEST_BATCH', 256))                 # per rank
    non_seq, seq, labels = create_sample_batch(cfg, B * world, (170, 170, 170), seed=77)
    to = lambda d, sl: {k: (v[sl].to(torch.bfloat16) if v.dim() == 3 else v[sl]).to(dev) for k, v in d.items()}
    sl = slice(rank * B, (rank + 1) * B)
    results = {}
    for overlap in (True, False):
        loss = train_step(model, grads, to(non_seq, sl), to(seq, sl), to(labels, sl), world, None, overlap)
        torch.cuda.synchronize()
        results[overlap] = (grads.flat.clone(), float(loss))
    dp_flat, dp_loss = results[True]
    loss_t = torch.tensor([dp_loss], device=dev)
    dist.all_reduce(loss_t, op=dist.ReduceOp.AVG)
    if rank == 0:
        whole = slice(0, B * world)
        names = {id(p): n for n, p in model.named_parameters()}

        def worst_rel(a_flat, b_flat):
            worst, worst_name = 0.0, ''
            for p, off in zip(grads.params, grads.offsets):
                a, b = a_flat[off:off + p.numel()].double(), b_flat[off:off + p.numel()].double()
                if float(b.norm()) == 0.0:
                    assert float(a.norm()) == 0.0, names[id(p)]
                    continue
                e = float((a - b).norm() / b.norm())
                if e > worst:
                    worst, worst_name = e, names[id(p)]
            return worst, worst_name

        one = []
        for _ in range(2):
            loss1 = train_step(model, grads, to(non_seq, whole), to(seq, whole), to(labels, whole), 1, None, False)
            torch.cuda.synchronize()
            one.append(grads.flat.clone())
        noise, noise_name = worst_rel(one[1], one[0])
        worst, worst_name = worst_rel(dp_flat, one[0])
        w2, w2_name = worst_rel(results[False][0], one[0])
        bound = min(max(3.0 * noise, 1e-3), 1.5e-2)
        out = {'test': 'dp_equivalence_2gpu', 'world': world, 'batch_per_rank': B, 'params': len(grads.params),
               'loss_mean_of_ranks': float(loss_t), 'loss_one_gpu_global_batch': float(loss1),
               'run_to_run_noise_one_gpu_rel_l2': noise, 'noisiest_tensor': noise_name,
               'dp_overlapped_vs_one_gpu_worst_rel_l2': worst, 'worst_tensor': worst_name,
               'dp_single_allreduce_vs_one_gpu_worst_rel_l2': w2, 'bound': bound,
               'pass': bool(worst <= bound and w2 <= bound and abs(float(loss_t) - float(loss1)) <= 1e-4 * max(1.0, abs(float(loss1))))}
        print(json.dumps(out), flush=True)
        os.makedirs('gpurun_out', exist_ok=True)
        json.dump(out, open('gpurun_out/r2_dp_equivalence_2gpu.json', 'w'), indent=1)
        assert out['pass'], out
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
