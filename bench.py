#!/usr/bin/env python
"""bench.py — headline benchmark of the OneTrans hot path on B200 (contract in the task statement).

Metric (BASELINE.json): OneTrans samples/sec, fwd+bwd bf16.  Workload at N GPUs (weak scaling, batch is the
shard): BASELINE config 2 — OneTrans-S, batch 2048 per GPU, 512 S + 32 NS tokens, pyramid pruning
``linear_to_ns`` down to the NS tokens — one step = zero grads + forward + BCE + backward (+ gradient
all-reduce over NCCL when N > 1).

  python bench.py --gpus 1 --steps 10 --warmup 3
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...      (CPU oracle port on the host cores; rank 0 only)

Prints ONE JSON line (rank 0).  `value` = device-resident inputs; `e2e` = same step fed from pinned host
buffers through the public module API with the loss read back every step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def workload(name: str):
    """BASELINE.json configs made concrete."""
    if name == 'c2':   # OneTrans-S training, B 2048, 512 S + 32 NS
        return dict(model='small', B=2048, seq_lens=(170, 170, 170), L_ns=32, schedule='linear_to_ns')
    if name == 'c3':   # OneTrans-L (d 384, 8 blocks), B 2048 per GPU
        return dict(model='default', B=2048, seq_lens=(170, 170, 170), L_ns=32, schedule='linear_to_ns')
    if name == 'c4':   # long sequence, halving schedule
        return dict(model='small', B=256, seq_lens=(672, 672, 672), L_ns=32, schedule='halving')
    if name == 'c5':   # cached inference: 1 user x 8192 candidates, NS-token-only queries (handled by run_c5)
        return dict(model='small', B=8192, seq_lens=(170, 170, 170), L_ns=32, schedule='linear_to_ns')
    if name == 'c1':   # the reference's CPU-runnable case
        return dict(model='small', B=32, seq_lens=(86, 84, 84), L_ns=16, schedule='reference_ratio')
    raise ValueError(name)


def fwd_flops_per_sample(d, F, H, n_layers, L0, keep_lens, L_s_events, n_ns_feat, L_ns, E=64):
    """Algorithmic forward FLOPs per sample: discarded work is not counted."""
    tot = 0.0
    cur = L0
    for keep in keep_lens:
        pairs = keep * cur - keep * (keep - 1) / 2.0
        tot += 4.0 * cur * d * d + 2.0 * keep * d * d + 4.0 * d * pairs + 2.0 * keep * d * d + 4.0 * keep * d * F
        cur = keep
    tot += 2.0 * L_s_events * E * d + 2.0 * n_ns_feat * d * L_ns
    return tot


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.gpu_index}', f'--query-gpu={q}', '--format=csv,noheader,nounits',
                                          '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def window(self, t0: float, t1: float):
        """Keep only the samples that arrived inside [t0, t1] (host clock around the synchronised timed region)."""
        self.t0, self.t1 = t0, t1

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        t0, t1 = getattr(self, 't0', None), getattr(self, 't1', None)
        rows = [r for (t, r) in self.rows if t0 is None or t0 <= t <= t1 + 0.15]
        if not rows:      # very short timed regions: fall back to every sample taken under load (warm-up included)
            rows = [r for (_, r) in self.rows]
        for r in rows:
            parts = [x.strip() for x in r.split(',')]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); smax.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(smax) if smax else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p['hbm_gbs'], tf_burst=p['bf16_tflops'], tf_sustained=p.get('bf16_tflops_sustained', p['bf16_tflops']), src='measured')
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src='fallback')


# ---------------------------------------------------------------------------------------------------
# CPU arm: the oracle port timed on the host cores (cpu_baseline, and --impl reference)
# ---------------------------------------------------------------------------------------------------

CPU_VARIANTS = {   # SURVEY.md §8d "CPU baseline beside it": the three structures of the same arithmetic
    'tail_only': dict(query_mode='tail_only', literal_loop=False),            # fair: same algorithmic FLOPs as the GPU path (D3)
    'literal_gather': dict(query_mode='literal_gather', literal_loop=False),  # reference FLOP structure: every query, then gather (OT/model.py:366-371)
    'literal_loop': dict(query_mode='literal_gather', literal_loop=True),     # reference dispatch structure: one tiny matmul per position (:84-88, 154-161)
}


def cpu_oracle_samples_per_sec(wl, sample_B: int, steps: int, warmup: int, dropout: float = 0.0, variant: str = 'tail_only'):
    from oracle import onetrans_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    ocfg = O.small_config(num_ns_tokens=wl['L_ns']) if wl['model'] == 'small' else O.default_config(num_ns_tokens=wl['L_ns'])
    ocfg.dropout_rate = dropout
    L0 = sum(wl['seq_lens']) + 2 + wl['L_ns']
    ocfg.pyramid_keep_lens = resolve_schedule(wl, ocfg.num_layers, L0)
    P = O.init_params(ocfg, seed=0)
    non_seq, seq, labels = O.synthetic_batch(ocfg, sample_B, wl['seq_lens'], seed=1234)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.loss_and_grads(P, ocfg, non_seq, seq, labels, training=dropout > 0, gen=torch.Generator().manual_seed(i), **CPU_VARIANTS[variant])
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    total = sum(times)
    return sample_B * len(times) / total, total / len(times) * 1e3, torch.get_num_threads()


def resolve_schedule(wl, n_layers, L0):
    from recommend_b200.schedule import keep_lens_halving, keep_lens_linear_to_ns, keep_lens_reference_ratio
    if wl['schedule'] == 'linear_to_ns':
        return keep_lens_linear_to_ns(L0, n_layers, wl['L_ns'])
    if wl['schedule'] == 'halving':
        return keep_lens_halving(L0, n_layers, wl['L_ns'])
    return keep_lens_reference_ratio(L0, n_layers, [0.5, 0.3, 0.2, 0.1, 0.05, 0.03, 0.02, 0.01])


def run_reference_arm(args, wl, rank):
    if rank != 0:
        return
    sample_B = args.cpu_sample_batch
    v, ms, cores = cpu_oracle_samples_per_sec(wl, sample_B, max(1, args.steps), max(0, args.warmup), args.dropout, args.cpu_variant)
    line = {
        'impl': 'reference', 'metric': 'OneTrans samples/sec (fwd+bwd bf16)', 'value': v, 'unit': 'samples/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        # the GPU arm's workload; each CPU step is a bounded sample of it: `cpu_step_batch` samples per step, not `global_batch`
        'config': dict(config_dict(args, wl, wl['B']), cpu_step_batch=sample_B,
                       note=f'reference arm: every timed step is {sample_B} samples of this workload on the host cores (a samples/s metric; '
                            f'the batch of {wl["B"]} per GPU is the GPU arm\'s)'),
        'cpu_baseline': {'value': v, 'unit': 'samples/s', 'cores': cores, 'kind': 'port',
                         'sample': f'oracle (PyTorch CPU fp32 restatement of OT/model.py, held at 1e-12 to vectors produced by the reference code itself over a TensorFlow-op shim; TensorFlow itself is not installable) fwd+BCE+bwd on {sample_B} '
                                   f'samples of the same workload per step; structure: {args.cpu_variant}'},
        'e2e': {'value': v, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


def config_dict(args, wl, B):
    return {'workload': f"{args.workload}: OneTrans-{'S' if wl['model'] == 'small' else 'L'} fwd+BCE+bwd, batch {B}/GPU, "
                        f"{sum(wl['seq_lens']) + 2} S + {wl['L_ns']} NS tokens, schedule {wl['schedule']}",
            'global_batch': B * args.gpus, 'seq_tokens': sum(wl['seq_lens']) + 2, 'ns_tokens': wl['L_ns'],
            'parallelism': f'dp{args.gpus}', 'dropout': args.dropout,
            'optimizer': 'clip_by_norm 90 + RMSprop inside the timed step' if args.optimizer else 'none (metric is fwd+bwd)',
            'grad_allreduce': 'none (1 GPU)' if args.gpus == 1 else (('after the backward' if args.no_overlap else 'per block, under the backward of the blocks below') +
                                                                     f", exchanged as {os.environ.get('OT_GRAD_REDUCE_DTYPE', 'fp32')} (fp32 accumulation buffer)"),
            'inputs': 'pre-embedded events bf16 [B, L_i, 64] x3, 11 fp32 scalars, 2 fp32 labels per sample (pinned host buffers in the e2e arm)',
            'l2_policy': 'activations per step (>20 GB) far exceed the 126 MB L2; no explicit flush'}


# ---------------------------------------------------------------------------------------------------
# BASELINE config 5: scoring with the cross-candidate K/V cache (north_star item 5; OT/model.py:95-98,120 repaired per D6)
# ---------------------------------------------------------------------------------------------------

def c5_cpu_candidates_per_sec(wl, n_cand: int, steps: int):
    """oracle two-stage scoring (user cache once, then ``n_cand`` candidates per step) on the host cores."""
    from oracle import onetrans_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    ocfg = O.small_config(num_ns_tokens=wl['L_ns'])
    ocfg.dropout_rate = 0.0
    L0 = sum(wl['seq_lens']) + 2 + wl['L_ns']
    ocfg.pyramid_keep_lens = resolve_schedule(wl, ocfg.num_layers, L0)
    P = O.init_params(ocfg, seed=0)
    non_seq, seq, _ = O.synthetic_batch(ocfg, n_cand, wl['seq_lens'], seed=1234)
    seq1 = {k: v[:1] for k, v in seq.items()}
    with torch.no_grad():
        cache = O.two_stage_user_cache(P, ocfg, seq1)
        O.two_stage_score(P, ocfg, cache, non_seq)
        t0 = time.perf_counter()
        for _ in range(steps):
            O.two_stage_score(P, ocfg, cache, non_seq)
        dt = (time.perf_counter() - t0) / steps
    return n_cand / dt, dt * 1e3, torch.get_num_threads()


def run_c5(args, wl, rank, local_rank, world):
    """`python bench.py --workload c5`: one user, 8192 candidates (sharded over the ranks when N > 1: strong scaling), the user's
    per-layer sequence-side K|V cached (stage 1) and every candidate's NS tokens scored against it (stage 2).  value = stage-2
    candidates/s with the cache and the candidate features resident; e2e = the request as a caller sees it: user sequences and
    candidate features from pinned host buffers, stage 1 + stage 2, probabilities back on the host."""
    C_total = wl['B']
    sample = min(args.cpu_sample_batch * 8, 256)
    if args.impl == 'reference':
        if rank != 0:
            return
        v, ms, cores = c5_cpu_candidates_per_sec(wl, sample, max(1, args.steps))
        print(json.dumps({'impl': 'reference', 'metric': 'OneTrans candidates/sec (cached inference)', 'value': v, 'unit': 'candidates/s',
                          'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True,
                          'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                          'config': {'workload': 'c5: OneTrans-S, 1 user x 8192 candidates, 512 S + 32 NS tokens, NS-token-only queries against the cached K|V',
                                     'cpu_step_candidates': sample},
                          'cpu_baseline': {'value': v, 'unit': 'candidates/s', 'cores': cores, 'kind': 'port',
                                           'sample': f'oracle two_stage_score (fp32) on {sample} candidates per step, user cache built once'},
                          'e2e': {'value': v, 'unit': 'candidates/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}), flush=True)
        return
    import torch.distributed as dist
    import recommend_b200 as R
    from recommend_b200 import _lib, ops
    from recommend_b200.data import create_sample_batch
    from recommend_b200.inference import score_candidates_sharded, shard_bounds
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    dev = torch.device('cuda', local_rank)
    cfg = R.get_model_config('small')
    cfg.num_ns_tokens, cfg.pyramid_schedule, cfg.dropout_rate = wl['L_ns'], wl['schedule'], 0.0
    torch.manual_seed(0)
    model = R.OneTransModel(cfg).to(dev).eval()
    non_seq, seq, _ = create_sample_batch(cfg, C_total, wl['seq_lens'], seed=1234)
    lo, hi = shard_bounds(C_total, world, rank)
    h_ns = {k: v[lo:hi].contiguous().pin_memory() for k, v in non_seq.items()}
    h_seq = {k: v[:1].to(torch.bfloat16).contiguous().pin_memory() for k, v in seq.items()}
    d_ns = {k: v.to(dev) for k, v in h_ns.items()}
    d_seq = {k: v.to(dev) for k, v in h_seq.items()}
    h2d = sum(v.numel() * v.element_size() for d_ in (h_ns, h_seq) for v in d_.values())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        for _ in range(max(3, args.warmup)):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            out = fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / n, out

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    with torch.no_grad():
        model.build_kv_cache(d_seq)
        l0 = _lib.launch_count
        t0 = time.perf_counter()
        ms2, _ = timed(lambda: model.score_candidates(d_ns), args.steps)
        sampler.window(t0, time.perf_counter())
        launches = (_lib.launch_count - l0) * args.steps // (args.steps + max(3, args.warmup))

        def request():
            seq_d = {k: v.to(dev, non_blocking=True) for k, v in h_seq.items()}
            ns_d = {k: v.to(dev, non_blocking=True) for k, v in h_ns.items()}
            model.build_kv_cache(seq_d)
            probs = model.score_candidates(ns_d)
            return {t: p.float().cpu() for t, p in probs.items()}        # device -> host read of the step's result
        ms_e2e, out = timed(request, args.steps)
        prof = ops.KernelProfiler()
        ops.set_profiler(prof)
        for _ in range(args.steps):
            model.score_candidates(d_ns)
        barrier()
        ops.set_profiler(None)
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        peaks = measured_peaks()
        summ = prof.summary()
        (name, tag), dmn = max(summ.items(), key=lambda kv: kv[1]['ms'])
        d2h = sum(v.numel() * 4 for v in out.values())
        line = {'metric': 'OneTrans candidates/sec (cached inference)', 'value': C_total / (ms2 * 1e-3), 'unit': 'candidates/s', 'n_gpus': world,
                'steps': args.steps, 'warmup': max(3, args.warmup), 'ms_per_step': ms2, 'higher_is_better': True, 'scaling': 'strong',
                'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
                'config': {'workload': 'c5: OneTrans-S, 1 user x 8192 candidates, 512 S + 32 NS tokens, NS-token-only queries against the cached per-layer K|V',
                           'candidates': C_total, 'parallelism': f'candidates sharded over {world} rank(s)', 'l2_policy': 'candidate activations per step (> 1 GB) exceed the 126 MB L2'},
                'clocks': clocks,
                'e2e': {'value': C_total / (ms_e2e * 1e-3), 'unit': 'candidates/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                        'step': 'host user sequences + candidate features -> stage 1 (cache) + stage 2 (scores) -> probabilities on the host'},
                'gpu_launches': launches,
                'roofline': roofline_of(name, tag, dmn, peaks),
                'kernels': [{'kernel': f'{n}[{tg}]', 'launches_per_step': d_['launches'] / args.steps, 'ms_per_step': d_['ms'] / args.steps}
                            for (n, tg), d_ in sorted(summ.items(), key=lambda kv: -kv[1]['ms'])[:10]]}
        if world == 1 and not args.no_cpu_baseline:
            v, ms, cores = c5_cpu_candidates_per_sec(wl, sample, 3)
            line['cpu_baseline'] = {'value': v, 'unit': 'candidates/s', 'cores': cores, 'kind': 'port',
                                    'sample': f'oracle two_stage_score (fp32) on {sample} candidates per step, 3 timed steps, user cache built once'}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='c2')
    ap.add_argument('--batch', type=int, default=0, help='override per-GPU batch')
    ap.add_argument('--cpu-sample-batch', type=int, default=32, help='samples per CPU-oracle step (cpu_baseline and --impl reference)')
    ap.add_argument('--cpu-variant', default='tail_only', choices=sorted(CPU_VARIANTS), help='--impl reference: tail_only (same FLOPs as the GPU path), '
                    'literal_gather (every query, then gather: the reference FLOP structure) or literal_loop (one matmul per position: its dispatch structure)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-kernel-profile', action='store_true')
    ap.add_argument('--dropout', type=float, default=0.1, help='training dropout rate (OT/config.py:50 default 0.1)')
    ap.add_argument('--no-overlap', action='store_true', help='N > 1: all-reduce the whole gradient buffer after the backward instead of per block under it')
    ap.add_argument('--optimizer', action='store_true', help='also run the clip + RMSprop update inside the step (OT/train.py:133-138)')
    args = ap.parse_args()

    wl = workload(args.workload)
    if args.batch:
        wl['B'] = args.batch
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))

    if args.workload == 'c5':
        run_c5(args, wl, rank, local_rank, world)
        return
    if args.impl == 'reference':
        run_reference_arm(args, wl, rank)
        return

    import torch.distributed as dist
    import recommend_b200 as R
    from recommend_b200 import _lib, ops
    from recommend_b200.train import FlatGradBuffer, ClipRMSprop, bce_loss, train_loop, train_step
    from recommend_b200.data import create_sample_batch

    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the product path has no CPU fallback')
    torch.cuda.set_device(local_rank)
    if world > 1:
        opts = dist.ProcessGroupNCCL.Options()
        opts.is_high_priority_stream = True       # the per-block gradient all-reduces slip in between the backward's kernels
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank), pg_options=opts)
    dev = torch.device('cuda', local_rank)

    cfg = R.get_model_config(wl['model'])
    cfg.num_ns_tokens = wl['L_ns']
    cfg.pyramid_schedule = wl['schedule']
    cfg.dropout_rate = args.dropout
    torch.manual_seed(0)
    model = R.OneTransModel(cfg).to(dev)
    grads = FlatGradBuffer(model.parameters())
    opt = ClipRMSprop.from_config(grads, cfg) if args.optimizer else None

    B = wl['B']
    non_seq, seq, labels = create_sample_batch(cfg, B, wl['seq_lens'], seed=1234 + rank)
    # host (pinned) copies for the e2e arm; device copies for the device-resident arm
    h_ns = {k: v.pin_memory() for k, v in non_seq.items()}
    h_seq = {k: v.to(torch.bfloat16).pin_memory() for k, v in seq.items()}
    h_lab = {k: v.pin_memory() for k, v in labels.items()}
    d_ns = {k: v.to(dev) for k, v in h_ns.items()}
    d_seq = {k: v.to(dev) for k, v in h_seq.items()}
    d_lab = {k: v.to(dev) for k, v in h_lab.items()}
    h2d_bytes = sum(v.numel() * v.element_size() for d_ in (h_ns, h_seq, h_lab) for v in d_.values())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        return train_step(model, grads, d_ns, d_seq, d_lab, world, opt, not args.no_overlap)

    def run_e2e(n_steps):
        """The user-facing loop (recommend_b200.train.train_loop): pinned host batches -> device (copies of step i+1
        overlap step i) -> train_step -> every step's loss read back on the host (one step behind the queue).
        Every step's inputs cross PCIe and every step's loss reaches the host inside the timed region."""
        losses = train_loop(model, grads, ((h_ns, h_seq, h_lab) for _ in range(n_steps)), world, opt, dev)
        assert len(losses) == n_steps
        return losses[-1]

    # ---- device-resident timing ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()       # nvidia-smi needs a few hundred ms to start: launch it before the warm-up, window it below
    for _ in range(max(3, args.warmup)):
        step_device()
    barrier()
    # which (kernel, shape) bucket dominates?  One fully instrumented step decides; the timed loop then carries CUDA
    # events on that bucket's launches only (two event records per launch on all ~170 launches cost ~3 % of a step)
    dom_key = None
    if not args.no_kernel_profile:
        probe = ops.KernelProfiler()
        ops.set_profiler(probe)
        step_device()
        barrier()
        ops.set_profiler(None)
        dom_key = max(probe.summary().items(), key=lambda kv: kv[1]['ms'])[0]
    dom_prof = ops.KernelProfiler(only=dom_key) if dom_key is not None else None
    launches0 = _lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_host0 = time.perf_counter()
    ops.set_profiler(dom_prof)
    e0.record()
    for _ in range(args.steps):
        loss = step_device()
    e1.record()
    barrier()
    ops.set_profiler(None)
    sampler.window(t_host0, time.perf_counter())
    launches = _lib.launch_count - launches0
    # per-kernel breakdown of every launch: a second pass over the same K steps, outside the headline loop
    prof = None if args.no_kernel_profile else ops.KernelProfiler()
    if prof is not None:
        ops.set_profiler(prof)
        for _ in range(args.steps):
            step_device()
        barrier()
        ops.set_profiler(None)
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = B * world / (ms_step * 1e-3)

    # ---- end-to-end timing (host buffers -> public API -> loss on host) ----
    run_e2e(2)
    barrier()
    e0.record()
    run_e2e(args.steps)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = B * world / (float(t.item()) / args.steps * 1e-3)

    ar_ms = None
    if world > 1:   # the one collective of the path, timed alone (device events, max over ranks)
        barrier()
        e0.record()
        for _ in range(args.steps):
            grads.all_reduce(world)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ar_ms = float(t.item()) / args.steps

    if rank == 0:
        peaks = measured_peaks()
        L0 = sum(wl['seq_lens']) + 2 + wl['L_ns']
        keep = R.resolve_keep_lens(cfg, L0)
        fwd_fl = fwd_flops_per_sample(cfg.hidden_dim, cfg.ffn_dim, cfg.num_heads, cfg.num_layers, L0, keep, sum(wl['seq_lens']),
                                      len(cfg.ns_features), wl['L_ns'])
        step_tflop = 3.0 * fwd_fl * B / 1e12
        line = {
            'metric': 'OneTrans samples/sec (fwd+bwd bf16)', 'value': value, 'unit': 'samples/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(3, args.warmup), 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'bf16', 'data': 'synthetic', 'config': config_dict(args, wl, B),
            'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': 'samples/s', 'h2d_bytes_per_step': h2d_bytes, 'd2h_bytes_per_step': 4},
            'gpu_launches': launches,
            'loss': float(loss),
            'grad_allreduce': None if ar_ms is None else {'ms': ar_ms, 'bytes': grads.flat.numel() * (2 if grads._low_precision() else 4),
                                                          'exchange_dtype': 'bf16' if grads._low_precision() else 'fp32',
                                                          'algbw_gbs': grads.flat.numel() * (2 if grads._low_precision() else 4) / (ar_ms * 1e-3) / 1e9,
                                                          'note': 'the collective alone (with its fp32 <-> bf16 conversion passes when the exchange is bf16)'},
            'model_flops': {'algorithmic_tflop_per_step_per_gpu': step_tflop, 'achieved_tflops_per_gpu': step_tflop / (ms_step * 1e-3),
                            'frac_of_bf16_sustained_peak': step_tflop / (ms_step * 1e-3) / peaks['tf_sustained'], 'peak_source': peaks['src']},
        }
        if prof is not None:
            summ = prof.summary()
            fam = {}
            for (name, tag), d_ in summ.items():
                f = fam.setdefault(name, dict(launches=0, ms=0.0, flops=0.0, bytes=0.0))
                for k in f:
                    f[k] += d_[k]
            kern_ms = sum(f['ms'] for f in fam.values())
            kernels = []
            for name, f in sorted(fam.items(), key=lambda kv: -kv[1]['ms']):
                kernels.append({'kernel': name, 'launches_per_step': f['launches'] / args.steps, 'ms_per_step': f['ms'] / args.steps,
                                'share': f['ms'] / kern_ms, 'tflops': f['flops'] / (f['ms'] * 1e-3) / 1e12, 'gbs': f['bytes'] / (f['ms'] * 1e-3) / 1e9})
            line['kernels'] = kernels
            try:   # per-shape detail for profiling notes (not part of the JSON line)
                os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
                det = [{'kernel': n, 'tag': tg, 'launches_per_step': d_['launches'] / args.steps, 'ms_per_step': d_['ms'] / args.steps,
                        'us_per_launch': d_['ms'] * 1e3 / d_['launches'], 'tflops': d_['flops'] / (d_['ms'] * 1e-3) / 1e12,
                        'gbs': d_['bytes'] / (d_['ms'] * 1e-3) / 1e9} for (n, tg), d_ in sorted(summ.items(), key=lambda kv: -kv[1]['ms'])]
                json.dump(det, open(os.path.join(ROOT, 'gpurun_out', 'kernels_detail.json'), 'w'), indent=1)
            except Exception:
                pass
            # dominant kernel: the (family, shape) bucket with the largest total time, timed inside the headline loop
            (name, tag), dmn = next(iter(dom_prof.summary().items()))
            line['roofline'] = roofline_of(name, tag, dmn, peaks)
            # the contraction kernels are judged on the tensor pipe, gather / norm kernels on HBM - one entry per family
            line['rooflines'] = [roofline_of(n, tg, d_, peaks, brief=True) for (n, tg), d_ in sorted(summ.items(), key=lambda kv: -kv[1]['ms'])[:12]]
            att = [(n, d_) for (n, tg), d_ in summ.items() if n in ('ot_attn_fwd', 'ot_attn_bwd')]
            if att:
                fl, ms_ = sum(d_['flops'] for _, d_ in att), sum(d_['ms'] for _, d_ in att)
                line['attn'] = {'ms_per_step': ms_ / args.steps, 'achieved_tflops': fl / (ms_ * 1e-3) / 1e12,
                                'frac_of_bf16_sustained_peak': fl / (ms_ * 1e-3) / 1e12 / peaks['tf_sustained'],
                                'tensor_pipe_util_pct': load_ncu_metric('attn_tensor_pipe_util_pct')}
        if world == 1 and not args.no_cpu_baseline:
            v, ms, cores = cpu_oracle_samples_per_sec(wl, args.cpu_sample_batch, 6, 1, args.dropout)
            line['cpu_baseline'] = {'value': v, 'unit': 'samples/s', 'cores': cores, 'kind': 'port',
                                    'sample': f'oracle (PyTorch CPU fp32 restatement of OT/model.py, held at 1e-12 to vectors produced by the reference code itself over a TensorFlow-op shim) fwd+BCE+bwd, 6 timed steps (1 warm-up) of '
                                              f'{args.cpu_sample_batch} samples of the same workload'}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


TENSOR_FAMILIES = ('ot_mixed_gemm', 'ot_ffn_fwd', 'ot_wgrad', 'ot_attn_fwd', 'ot_attn_bwd', 'ot_attn_ns_cached_fwd')


def roofline_of(name, tag, d, peaks, brief=False):
    """Roofline entry of one (kernel family, shape) bucket from CUDA-event time and algorithmic work.  SURVEY.md 8(d): the
    dense contractions (grouped GEMMs, fused FFN, weight gradients, attention) are bound by the TENSOR pipe and reported against the
    measured sustained bf16 peak; tokenizer / norm / optimizer / head kernels by HBM against the measured copy bandwidth.  The other
    side is always given too (`hbm` with the op-minimal bytes, see ops.KernelProfiler), so that a reader can see which wall is closer."""
    secs = d['ms'] * 1e-3
    tf = d['flops'] / secs / 1e12
    gbs_min = d.get('min_bytes', d['bytes']) / secs / 1e9
    if name in TENSOR_FAMILIES:
        roof = {'bound': 'tensor', 'achieved': tf, 'peak': peaks['tf_sustained'], 'unit': 'TFLOP/s', 'frac': tf / peaks['tf_sustained'],
                'hbm': {'achieved': gbs_min, 'peak': peaks['hbm'], 'unit': 'GB/s', 'frac': gbs_min / peaks['hbm'],
                        'bytes': 'op-minimal: operands and one output, no second outputs'}}
    else:
        gbs = d['bytes'] / secs / 1e9
        roof = {'bound': 'hbm', 'achieved': gbs, 'peak': peaks['hbm'], 'unit': 'GB/s', 'frac': gbs / peaks['hbm']}
    roof['kernel'] = f'{name}[{tag}]'
    roof['ms_per_launch'] = d['ms'] / d['launches']
    if brief:
        return roof
    tr, tr_src = load_ncu_traffic(name, tag, d['bytes'] / d['launches'])
    roof.update({'launches': d['launches'], 'avg_launch_us': d['ms'] * 1e3 / d['launches'],
                 'algorithmic_bytes_per_launch': d['bytes'] / d['launches'], 'algorithmic_flops_per_launch': d['flops'] / d['launches'],
                 'peak_source': peaks['src'] + (' (sustained bf16, MEASURED_PEAKS.json)' if roof['bound'] == 'tensor' else ' (copy bandwidth, MEASURED_PEAKS.json)'),
                 'traffic': tr, 'traffic_source': tr_src,
                 'timed': 'CUDA events around every launch of this bucket inside the timed region; kernels[] / rooflines[] come from a second, fully instrumented pass of the same steps'})
    return roof


def load_ncu_metric(key):
    """A number copied from a committed ncu capture of this build (profiles/r2_ncu_metrics.json), with its provenance; None if absent."""
    path = os.path.join(ROOT, 'profiles', 'r2_ncu_metrics.json')
    try:
        return json.load(open(path)).get(key)
    except Exception:
        return None


def load_ncu_traffic(family, tag=None, algorithmic_bytes_per_launch=None):
    """(DRAM bytes per launch, provenance) of the dominant kernel from the committed ncu capture (profiles/ncu_traffic.json), or
    (None, why).  The capture holds layer-0 shapes; the bench bucket averages one shape family over the layers, so the measured
    traffic / algorithmic ratio of the captured launch is applied to the bucket's algorithmic bytes per launch and labelled so."""
    path = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
    if not os.path.exists(path):
        return None, 'no committed capture'
    try:
        d = json.load(open(path))
        e = d.get(f'{family}[{tag}]')
        if isinstance(e, dict) and e.get('traffic_over_algorithmic') and algorithmic_bytes_per_launch:
            return (e['traffic_over_algorithmic'] * algorithmic_bytes_per_launch,
                    f"extrapolated: dram__bytes_read.sum + dram__bytes_write.sum of the layer-0 launch of this bucket in {e.get('capture', 'profiles/')} "
                    f"({e['dram_bytes'] / 1e9:.3f} GB measured / {e['algorithmic_bytes'] / 1e9:.3f} GB algorithmic = {e['traffic_over_algorithmic']:.3f}) "
                    f"x this bucket's algorithmic bytes per launch")
        return None, f'no capture of {family}[{tag}]'
    except Exception as ex:
        return None, f'unreadable capture: {ex}'


if __name__ == '__main__':
    main()
