"""Data-parallel host logic on CPU: world_size-2 gloo run of FlatGradBuffer.all_reduce (the only exchange on
the path).  Each rank computes oracle gradients on its half of a global batch; after the
all-reduce every rank must hold the gradients of the full batch (T10, DP half)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import onetrans_oracle as O
from recommend_b200.train import FlatGradBuffer, bce_loss


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _cfg():
    return O.OracleConfig(hidden_dim=32, num_layers=2, num_heads=2, ffn_dim=64, num_ns_tokens=3, dropout_rate=0.0)


def _worker(rank, world, port, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(1)
    cfg = _cfg()
    P = O.init_params(cfg, seed=0)
    non_seq, seq, labels = O.synthetic_batch(cfg, 8, (5, 4, 3), seed=7)   # the GLOBAL batch, identical on every rank
    per = 8 // world
    sl = slice(rank * per, (rank + 1) * per)                               # rank r takes rows [r*B_local, (r+1)*B_local)
    shard = lambda d: {k: v[sl] for k, v in d.items()}
    _, grads, _ = O.loss_and_grads(P, cfg, shard(non_seq), shard(seq), shard(labels))
    names = sorted(P)
    params = [torch.nn.Parameter(P[n].clone()) for n in names]
    buf = FlatGradBuffer(params)
    for p, n in zip(params, names):
        p.grad.copy_(grads[n])
    buf.all_reduce(world, bucket_bytes=4096)                               # several buckets
    torch.save({n: p.grad.clone() for p, n in zip(params, names)}, os.path.join(out_dir, f'rank{rank}.pt'))
    dist.destroy_process_group()


def _worker_ranges(rank, world, port, out_dir):
    """Overlapped reduction API: two 'blocks' reduced early by range, the rest swept up by finish_overlapped_reduce."""
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    g = torch.Generator().manual_seed(100 + rank)
    params = [torch.nn.Parameter(torch.zeros(n)) for n in (5, 1500, 7, 2048, 3)]
    buf = FlatGradBuffer(params)
    for p in params:
        p.grad.copy_(torch.randn(p.shape, generator=g))
    local = [p.grad.clone() for p in params]
    buf.begin_overlapped_reduce(world)
    buf.reduce_range_async(buf.range_of(params[3:4]))          # "block 1" finishes its backward first
    buf.reduce_range_async(buf.range_of(params[1:3]))          # then "block 0"
    buf.finish_overlapped_reduce()                             # params 0 and 4 were in no bucket
    torch.save({'local': local, 'reduced': [p.grad.clone() for p in params]}, os.path.join(out_dir, f'ranges{rank}.pt'))
    dist.destroy_process_group()


def test_overlapped_range_reduce_covers_every_parameter_once(tmp_path):
    world = 2
    mp.spawn(_worker_ranges, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(os.path.join(tmp_path, f'ranges{i}.pt')) for i in range(world)]
    for i in range(5):
        want = (r[0]['local'][i] + r[1]['local'][i]) / world
        assert torch.allclose(r[0]['reduced'][i], want, rtol=1e-6, atol=1e-7), i
        assert torch.equal(r[0]['reduced'][i], r[1]['reduced'][i]), i


def test_allreduced_shard_grads_equal_full_batch_grads(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    cfg = _cfg()
    P = O.init_params(cfg, seed=0)
    non_seq, seq, labels = O.synthetic_batch(cfg, 8, (5, 4, 3), seed=7)
    _, full, _ = O.loss_and_grads(P, cfg, non_seq, seq, labels)
    r0 = torch.load(os.path.join(tmp_path, 'rank0.pt'))
    r1 = torch.load(os.path.join(tmp_path, 'rank1.pt'))
    for n in full:
        assert torch.equal(r0[n], r1[n]), n                                # every rank holds the same reduced gradient
        assert torch.allclose(r0[n], full[n], rtol=1e-4, atol=1e-6), n


def test_flat_grad_buffer_views_and_zero():
    ps = [torch.nn.Parameter(torch.randn(3, 4)), torch.nn.Parameter(torch.randn(5))]
    buf = FlatGradBuffer(ps)
    A = FlatGradBuffer.ALIGN                      # every tensor starts on an optimizer-chunk boundary
    assert buf.offsets == [0, A] and buf.flat.numel() == 2 * A
    ps[0].grad.add_(1.0)
    ps[1].grad.add_(2.0)
    assert buf.flat[:12].eq(1).all() and buf.flat[A:A + 5].eq(2).all()
    assert buf.flat[12:A].eq(0).all() and buf.flat[A + 5:].eq(0).all()
    buf.zero()
    assert ps[0].grad.abs().sum() == 0 and ps[1].grad.abs().sum() == 0
    buf.all_reduce(1)   # world 1: no-op, no process group needed


def test_bce_matches_oracle():
    torch.manual_seed(0)
    p = {'ctr': torch.rand(6, 1), 'cvr': torch.rand(6, 1)}
    y = {'ctr': (torch.rand(6, 1) < 0.5).float(), 'cvr': (torch.rand(6, 1) < 0.5).float()}
    assert torch.allclose(bce_loss(p, y, ['ctr', 'cvr']), O.bce_loss(p, y, ['ctr', 'cvr']))


# ---- sharded candidate scoring (BASELINE config 5 on several GPUs): the host side of the one all-gather ----------
def _worker_gather(rank, world, port, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from recommend_b200.inference import shard_bounds, gather_shards
    res = {}
    for n in (7, 8, 1, 2):                                    # ragged shards, even shards, an empty shard on the last rank
        scores = torch.arange(2 * n, dtype=torch.float32).reshape(2, n) * 0.5 + 1.0      # the [T, n] table every rank should end with
        s, e = shard_bounds(n, world, rank)
        res[n] = gather_shards(scores[:, s:e].contiguous(), n, world, rank)
    torch.save(res, os.path.join(out_dir, f'gather{rank}.pt'))
    dist.destroy_process_group()


def test_sharded_candidate_scores_gather_in_order(tmp_path):
    from recommend_b200.inference import shard_bounds
    for n, w in ((8192, 8), (7, 2), (5, 8), (0, 4)):
        b = [shard_bounds(n, w, r) for r in range(w)]
        assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        assert max(e - s for s, e in b) - min(e - s for s, e in b) <= 1
    world = 2
    mp.spawn(_worker_gather, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(os.path.join(tmp_path, f'gather{i}.pt')) for i in range(world)]
    for n in (7, 8, 1, 2):
        want = torch.arange(2 * n, dtype=torch.float32).reshape(2, n) * 0.5 + 1.0
        assert torch.equal(r[0][n], want) and torch.equal(r[1][n], want), n


# ---- data-parallel evaluation: metric states are counts, so the sum over ranks is the state of the whole dataset ---------------
def _metric_state(y, p, nt):
    """What ot_metrics_update leaves in the state, written with the oracle (layout of include/onetrans_b200.h)."""
    import numpy as np
    from oracle import metrics_oracle as M
    st = torch.zeros(2, 2 * nt + 8, dtype=torch.int64)
    for t in range(2):
        pos, neg = M.keras_auc_state(y[t], p[t], nt)
        c = M.confusion_counts(y[t], p[t])
        st[t, :nt], st[t, nt:2 * nt] = torch.from_numpy(pos), torch.from_numpy(neg)
        st[t, 2 * nt:2 * nt + 5] = torch.tensor([c['tp'], c['fp'], c['tn'], c['fn'], c['count']])
        st[t, 2 * nt + 6] = torch.tensor([M.binary_crossentropy(y[t], p[t]) * c['count']], dtype=torch.float64).view(torch.int64)[0]
    return st


def _worker_metrics(rank, world, port, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import numpy as np
    from recommend_b200.metrics import merge_states_
    rng = np.random.default_rng(0)                                   # the GLOBAL evaluation set, identical on every rank
    p = rng.random((2, 1001)).astype(np.float32)
    y = (rng.random((2, 1001)) < p).astype(np.float32)
    s, e = (0, 400) if rank == 0 else (400, 1001)                    # uneven shards
    st = _metric_state(y[:, s:e], p[:, s:e], 200)
    merge_states_(st, 200)
    torch.save(st, os.path.join(out_dir, f'metrics{rank}.pt'))
    dist.destroy_process_group()


def test_metric_states_sum_over_ranks_to_the_whole_dataset(tmp_path):
    import numpy as np
    world = 2
    mp.spawn(_worker_metrics, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(0)
    p = rng.random((2, 1001)).astype(np.float32)
    y = (rng.random((2, 1001)) < p).astype(np.float32)
    want = _metric_state(y, p, 200)
    got = [torch.load(os.path.join(tmp_path, f'metrics{i}.pt')) for i in range(world)]
    k = 2 * 200 + 6
    for g in got:
        ints = [c for c in range(g.shape[1]) if c != k]
        assert torch.equal(g[:, ints], want[:, ints])                                     # every count exact
        assert torch.allclose(g[:, k].clone().view(torch.float64), want[:, k].clone().view(torch.float64), rtol=1e-12)
    assert torch.equal(got[0], got[1])
