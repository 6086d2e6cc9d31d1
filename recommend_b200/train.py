"""Training-step host logic for the OneTrans path: BCE loss as the reference builds it
(OT/train.py:78-93, 124-128), a flat fp32 gradient buffer shared by every parameter, and the
data-parallel gradient all-reduce (PAPER:190; the reference code has no distributed strategy).

One process per GPU; ``torch.distributed`` (NCCL over NVLink/NVSwitch) is the plumbing."""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Tuple

import os
import torch
import torch.distributed as dist


def bce_loss(probs: Dict[str, torch.Tensor], labels: Dict[str, torch.Tensor], tasks: Iterable[str]) -> torch.Tensor:
    """Sum over tasks of Keras ``BinaryCrossentropy(from_logits=False, SUM_OVER_BATCH_SIZE)``
    (OT/train.py:84-87, 124-128): clip p to [1e-7, 1-1e-7], -[y log(p+1e-7) + (1-y) log(1-p+1e-7)], batch mean."""
    eps = 1e-7
    total = None
    for t in tasks:
        if t in probs and t in labels:
            p = probs[t].float().clamp(eps, 1.0 - eps)
            y = labels[t].to(p.dtype)
            l = (-(y * torch.log(p + eps) + (1.0 - y) * torch.log(1.0 - p + eps))).mean()
            total = l if total is None else total + l
    return total


class FlatGradBuffer:
    """One contiguous fp32 buffer holding the gradient of every parameter; ``p.grad`` are views into it.
    The weight-gradient kernels accumulate into the views directly, ``zero()`` is one memset and the
    data-parallel reduction is one (bucketed) all-reduce of the buffer."""

    ALIGN = 1024  # elements; == OT_OPT_CHUNK so an optimizer chunk never straddles two tensors

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.offsets: List[int] = []
        off = 0
        for p in self.params:
            self.offsets.append(off)
            off += -(-p.numel() // self.ALIGN) * self.ALIGN
        dev = self.params[0].device
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        for p, o in zip(self.params, self.offsets):
            p.grad = self.flat[o:o + p.numel()].view(p.shape)
        # Data-parallel exchange format (SURVEY 7.2 "measure"): 'bf16' sends a bf16 copy of each bucket through the collective and
        # writes the average back into the fp32 buffer - half the bytes on NVLink and half the time NCCL's CTAs hold SMs that the
        # one-CTA-per-SM persistent kernels of the backward want; the 2^-9 rounding of the AVERAGED gradient is far inside the
        # run-to-run noise of the gradients themselves (fp32 atomics, bf16 dQ reductions: 0.7e-2, profiles/r2_dp_equivalence_2gpu.json).
        # Measured on 2 x B200 (profiles/README.md): 38.9 ms/step with the bf16 exchange against 38.1 with fp32 - at two GPUs the
        # collective alone takes 1.2 ms either way and the two conversion passes cost more than the halved payload saves, so fp32
        # stays the default; OT_GRAD_REDUCE_DTYPE=bf16 selects the bf16 exchange (NCCL only: gloo has no bf16 AVG).
        self.reduce_dtype = os.environ.get('OT_GRAD_REDUCE_DTYPE', 'fp32')
        self._lowp: Optional[torch.Tensor] = None

    def zero(self) -> None:
        self.flat.zero_()

    def all_reduce(self, world_size: int, bucket_bytes: int = 0) -> None:
        """Average over ranks (PAPER:190 data parallelism; the reference code is single-device).  On NCCL the
        averaging rides in the collective (``ReduceOp.AVG``), so the 4 B/parameter scaling pass disappears; one call
        over the whole buffer unless ``bucket_bytes`` asks for pieces."""
        if world_size <= 1:
            return
        nccl = dist.get_backend() == 'nccl'
        op = dist.ReduceOp.AVG if nccl else dist.ReduceOp.SUM
        n_per = self.flat.numel() if bucket_bytes <= 0 else max(1, bucket_bytes // 4)
        for s in range(0, self.flat.numel(), n_per):
            if self._low_precision():
                lp = self._lowp_view(s, min(s + n_per, self.flat.numel()))
                lp.copy_(self.flat[s:s + n_per])
                dist.all_reduce(lp, op=op)
                self.flat[s:s + n_per].copy_(lp)
            else:
                dist.all_reduce(self.flat[s:s + n_per], op=op)
        if not nccl:
            self.flat.mul_(1.0 / world_size)

    def _low_precision(self) -> bool:
        return self.reduce_dtype == 'bf16' and dist.is_initialized() and dist.get_backend() == 'nccl'

    def _lowp_view(self, s: int, e: int) -> torch.Tensor:
        if self._lowp is None:
            self._lowp = torch.empty(self.flat.numel(), dtype=torch.bfloat16, device=self.flat.device)
        return self._lowp[s:e]

    # ---- overlapped reduction: one asynchronous all-reduce per block, issued as soon as the block's backward is enqueued ----
    def range_of(self, params: Iterable[torch.nn.Parameter]):
        """[start, end) of the flat buffer covering ``params`` (which must be consecutive in the buffer)."""
        index = {id(p): i for i, p in enumerate(self.params)}
        idx = sorted(index[id(p)] for p in params if id(p) in index)
        if not idx:
            return None
        assert idx == list(range(idx[0], idx[-1] + 1)), 'parameters of one bucket must be consecutive in the flat buffer'
        last = self.params[idx[-1]]
        end = self.offsets[idx[-1]] + -(-last.numel() // self.ALIGN) * self.ALIGN
        return self.offsets[idx[0]], end

    def begin_overlapped_reduce(self, world_size: int) -> None:
        self._works, self._done, self._world = [], [], world_size
        self._copy_back = []

    def reduce_range_async(self, rng) -> None:
        """Enqueue the all-reduce of flat[start:end] behind everything already on the current stream; later kernels of
        the current stream run concurrently with it (the persistent kernels schedule their tiles dynamically, so SMs
        that the collective occupies cost throughput, not stragglers)."""
        if rng is None or self._world <= 1:
            return
        nccl = dist.get_backend() == 'nccl'
        if self._low_precision():
            lp = self._lowp_view(rng[0], rng[1])
            lp.copy_(self.flat[rng[0]:rng[1]])          # fp32 -> bf16 on the current stream, behind the kernels that produced the bucket
            self._works.append(dist.all_reduce(lp, op=dist.ReduceOp.AVG, async_op=True))
            self._copy_back.append(rng)
        else:
            self._works.append(dist.all_reduce(self.flat[rng[0]:rng[1]], op=dist.ReduceOp.AVG if nccl else dist.ReduceOp.SUM, async_op=True))
        self._done.append(rng)

    def finish_overlapped_reduce(self) -> None:
        """Reduce whatever no bucket covered, then make the current stream wait for every outstanding collective."""
        if self._world > 1:
            covered, pos = sorted(self._done), 0
            for s, e in covered + [(self.flat.numel(), self.flat.numel())]:
                if s > pos:
                    self.reduce_range_async((pos, s))
                pos = max(pos, e)
            for w in self._works:
                w.wait()
            for s, e in self._copy_back:                 # averaged bf16 buckets back into the fp32 gradient buffer
                self.flat[s:e].copy_(self._lowp[s:e])
            if dist.get_backend() != 'nccl':
                self.flat.mul_(1.0 / self._world)
        self._works, self._done, self._copy_back = [], [], []


class DevicePrefetcher:
    """Host -> device input pipeline for the train loop (the role of ``tf.data`` prefetching in OT/data_loader.py:
    195-222): iterates ``(non_seq, seq, labels)`` dictionaries of pinned host tensors and hands out device copies,
    always keeping the NEXT batch's copies in flight on a side stream so that they overlap the current step.
    Two sets of device buffers are allocated once and reused alternately (no allocator traffic in steady state)."""

    def __init__(self, batches, device: torch.device):
        self.it = iter(batches)
        self.device = device
        self.copy_stream = torch.cuda.Stream(device=device)
        self.slots = [None, None]
        self.free_ev = [None, None]      # compute-stream event after the last step that read the slot
        self.n = 0
        self._next = None
        self._preload()

    def _preload(self) -> None:
        try:
            host = next(self.it)
        except StopIteration:
            self._next = None
            return
        k = self.n % 2
        self.n += 1
        if self.slots[k] is None or any(d[key].shape != h[key].shape or d[key].dtype != h[key].dtype
                                        for d, h in zip(self.slots[k], host) for key in h):
            self.slots[k] = tuple({key: torch.empty(v.shape, dtype=v.dtype, device=self.device) for key, v in h.items()} for h in host)
        with torch.cuda.stream(self.copy_stream):
            if self.free_ev[k] is not None:
                self.copy_stream.wait_event(self.free_ev[k])
            for d, h in zip(self.slots[k], host):
                for key, v in h.items():
                    d[key].copy_(v, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self._next = (k, ev)

    def __iter__(self):
        return self

    def __next__(self):
        if self._next is None:
            raise StopIteration
        k, ev = self._next
        cur = torch.cuda.current_stream(self.device)
        # everything enqueued so far (the previous step) read the OTHER slot: it may be overwritten once that is done
        done = torch.cuda.Event()
        done.record(cur)
        self.free_ev[1 - k] = done
        cur.wait_event(ev)
        self._preload()
        return self.slots[k]


def train_loop(model, grads: FlatGradBuffer, host_batches, world_size: int = 1, optimizer: Optional['ClipRMSprop'] = None,
               device: Optional[torch.device] = None) -> List[float]:
    """The epoch loop of OT/train.py:186-232 around ``train_step``: pinned host batches in, per-step losses out.
    Inputs of step i+1 are copied while step i runs (DevicePrefetcher); the loss of step i travels to pinned host
    memory asynchronously and is read after step i+1 has been enqueued, so the host never drains the GPU queue
    between steps.  Returns one Python float per step."""
    device = device if device is not None else grads.flat.device
    losses: List[float] = []
    pending = None      # (pinned host scalar, event)
    for non_seq, seq, labels in DevicePrefetcher(host_batches, device):
        loss = train_step(model, grads, non_seq, seq, labels, world_size, optimizer)
        h = torch.empty((), dtype=torch.float32, pin_memory=True)
        h.copy_(loss, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        if pending is not None:
            pending[1].synchronize()
            losses.append(float(pending[0]))
        pending = (h, ev)
    if pending is not None:
        pending[1].synchronize()
        losses.append(float(pending[0]))
    return losses


def broadcast_parameters(model, world_size: int, src: int = 0) -> None:
    """Data parallel keeps replicas identical only if they start identical: copy rank ``src``'s parameters to every rank (one
    broadcast per tensor, once after construction / loading).  No-op for one rank or without an initialised process group."""
    if world_size <= 1 or not dist.is_available() or not dist.is_initialized():
        return
    with torch.no_grad():
        for p in model.parameters():
            dist.broadcast(p.data, src=src)
            torch.autograd.graph.increment_version(p)     # bf16 compute copies refresh on the next forward


class ClipRMSprop:
    """The dense-parameter update of the reference train step (OT/train.py:131-138): per-tensor
    ``tf.clip_by_norm(g, gradient_clip_norm)`` followed by Keras ``RMSprop(learning_rate, rho=0.9, momentum,
    epsilon=1e-7)`` (OT/train.py:65-70; lr 0.005, momentum 0.99999, clip 90 from OT/config.py:39-52), as two
    streaming kernels over the flat gradient buffer (``ot_clip_rmsprop_step``).  fp32 masters are updated in place and
    their version counters bumped so the bf16 compute copies refresh on the next forward."""

    def __init__(self, grads: FlatGradBuffer, lr: float = 0.005, rho: float = 0.9, momentum: float = 0.0,
                 eps: float = 1e-7, clip_norm: float = 0.0, per_variable: bool = True):
        from . import ops
        self._ops = ops
        self.grads = grads
        self.lr, self.rho, self.momentum, self.eps, self.clip_norm = float(lr), float(rho), float(momentum), float(eps), float(clip_norm)
        dev = grads.flat.device
        if dev.type != 'cuda':
            raise RuntimeError('ClipRMSprop runs on CUDA tensors only (no CPU fallback)')
        for p in grads.params:
            if p.dtype != torch.float32 or not p.is_contiguous() or p.data_ptr() % 16:
                raise RuntimeError('ClipRMSprop: parameters must be contiguous, 16-byte aligned fp32 CUDA tensors')
        self.rms = torch.zeros_like(grads.flat)
        self.mom = torch.zeros_like(grads.flat) if self.momentum != 0.0 else None
        # Clip granularity = the reference's Keras variables (OT/train.py:135 clips each entry of model.trainable_variables):
        # the packed tensors (Wqkv [G, d, 3d], W1 / b1 / W2 / b2 [G, ...]) carry ``_ot_clip_layout`` and split into one clip
        # slot per (weight group, q | k | v part); ``per_variable=False`` clips each packed tensor as a whole instead.
        slots, base = [], 0
        self.slot_names: List[Tuple[int, int]] = []     # (parameter index, first slot) for grad_norms()
        for i, p in enumerate(grads.params):
            n = max(p.numel(), 1)
            outer, row, part = getattr(p, '_ot_clip_layout', (n, n, n)) if per_variable else (n, n, n)
            if (outer, row, part) != (n, n, n) and (part % 4 or row % part or outer % row or n % outer):
                raise RuntimeError(f'ClipRMSprop: unsupported clip layout {(outer, row, part)} for a tensor of {n} elements')
            slots.append((base, outer, row, part))
            self.slot_names.append((i, base))
            base += (n // outer) * (row // part)
        self.n_slots = base
        self._seg_slot = torch.tensor(slots, dtype=torch.int64, device=dev).contiguous() if base != len(grads.params) else None
        self.sqnorm = torch.zeros(self.n_slots, dtype=torch.float32, device=dev)
        self._seg_off = torch.tensor(grads.offsets + [grads.flat.numel()], dtype=torch.int64, device=dev)
        self._seg_numel = torch.tensor([p.numel() for p in grads.params], dtype=torch.int64, device=dev)
        self._ptrs = None
        self._ptr_key = None

    @classmethod
    def from_config(cls, grads: FlatGradBuffer, config) -> 'ClipRMSprop':
        oc = config.optimizer_config
        return cls(grads, lr=oc.get('dense_lr', 0.005), rho=oc.get('rho', 0.9), momentum=oc.get('momentum', 0.0),
                   eps=oc.get('epsilon', 1e-7), clip_norm=getattr(config, 'gradient_clip_norm', 0.0))

    def _param_table(self) -> torch.Tensor:
        key = tuple(p.data_ptr() for p in self.grads.params)
        if key != self._ptr_key:
            self._ptrs = torch.tensor(key, dtype=torch.int64, device=self.grads.flat.device)
            self._ptr_key = key
        return self._ptrs

    @torch.no_grad()
    def step(self, grad_scale: float = 1.0, zero_grad: bool = False) -> None:
        self._ops.clip_rmsprop_step(self._param_table(), self._seg_off, self._seg_numel, self.grads.flat, self.rms, self.mom,
                                    self.sqnorm, lr=self.lr, rho=self.rho, momentum=self.momentum, eps=self.eps,
                                    clip_norm=self.clip_norm, grad_scale=grad_scale, zero_grad=zero_grad,
                                    seg_slot=self._seg_slot, n_slots=self.n_slots)
        for p in self.grads.params:
            torch.autograd.graph.increment_version(p)

    def grad_norms(self) -> torch.Tensor:
        """Gradient L2 norms per clip slot (Keras variable; ``slot_names`` maps parameters to their first slot) seen by the last
        ``step`` (valid when clip_norm > 0)."""
        return self.sqnorm.sqrt()


def train_step(model, grads: FlatGradBuffer, non_seq, seq, labels, world_size: int = 1,
               optimizer: Optional[ClipRMSprop] = None, overlap_reduce: bool = True) -> torch.Tensor:
    """forward + BCE + backward (+ gradient all-reduce): the "fwd+bwd" of the headline metric
    (OT/train.py:116-131); with ``optimizer`` also the clip + RMSprop update (OT/train.py:133-138).
    With more than one rank the gradients of block l are all-reduced while the backward of blocks l-1 ... 0 runs
    (``overlap_reduce``; PAPER:190 data parallelism).  Returns the detached loss tensor (no host sync)."""
    from . import engine
    grads.zero()
    if hasattr(model, 'forward_with_loss'):
        loss, _ = model.forward_with_loss(non_seq, seq, labels, training=True)      # BCE inside the head kernel
    else:
        loss = bce_loss(model(non_seq, seq, training=True), labels, model.config.tasks)
    if world_size > 1 and overlap_reduce and hasattr(model, 'blocks'):
        if getattr(grads, '_block_ranges', None) is None:
            # two slices per block, in the order their gradients become final: the FFN weights (about 3/4 of a block),
            # then the norms and the attention weights
            grads._ffn_ranges = {id(b): grads.range_of(list(b.ffn.parameters())) for b in model.blocks}
            grads._block_ranges = {id(b): grads.range_of([p for p in b.parameters() if all(p is not q for q in b.ffn.parameters())])
                                   for b in model.blocks}
        grads.begin_overlapped_reduce(world_size)
        engine.after_ffn_backward = lambda blk: grads.reduce_range_async(grads._ffn_ranges.get(id(blk)))
        engine.after_block_backward = lambda blk: grads.reduce_range_async(grads._block_ranges.get(id(blk)))
        try:
            loss.backward()
        finally:
            engine.after_ffn_backward = engine.after_block_backward = None
        grads.finish_overlapped_reduce()
    else:
        loss.backward()
        grads.all_reduce(world_size)
    if optimizer is not None:
        optimizer.step()
    return loss.detach()


class OneTransTrainer:
    """The reference's trainer object (OT/train.py:19-338) on the sm_100a path: same constructor and method names
    (``train_step`` / ``val_step`` / ``train`` / ``save_model`` / ``load_model``), same history dictionary, same early-stopping and
    checkpoint cadence.  What differs, and why:

    * batches are ``(non_seq_features, seq_features, labels)`` triples of host or device tensors; the model is called with two
      arguments (the reference passes one tuple, which its own ``call`` rejects - SURVEY.md D7);
    * the per-tensor clip reads ``config.gradient_clip_norm`` (the reference reads a non-existent ``gradient_clip``, D8);
    * dense optimizer: RMSprop only (``ot_clip_rmsprop_step``); the reference's ``'adam'`` branch (:56-62) is not on the path that
      OT/config.py:39-47 selects and raises here rather than silently running something else;
    * metrics are the streaming kernels of ``recommend_b200.metrics`` (all tasks in one pass);
    * weights go to ``model_weights.npz`` in Keras weight order (``state.save_weights``; ``.h5`` needs h5py)."""

    def __init__(self, config, model_dir: str = './models', device: str = 'cuda', world_size: int = 1):
        from pathlib import Path
        from .model import OneTransModel
        from .metrics import BinaryTaskMetrics
        self.config = config
        self.model_dir = Path(model_dir)
        self.model_dir.mkdir(parents=True, exist_ok=True)
        self.device = torch.device(device)
        self.world_size = world_size
        self.model = OneTransModel(config).to(self.device)                       # OT/train.py:28
        broadcast_parameters(self.model, world_size)                             # replicas start from rank 0's initialisation
        self.grads = FlatGradBuffer(self.model.parameters())
        self.optimizer = self._create_optimizer()                                # :31
        self.train_metrics = BinaryTaskMetrics(config.tasks, self.device)        # :37-38
        self.val_metrics = BinaryTaskMetrics(config.tasks, self.device)
        self.history = {'train_loss': [], 'val_loss': [], 'train_metrics': {}, 'val_metrics': {}}      # :41-46

    def _create_optimizer(self) -> ClipRMSprop:
        kind = self.config.optimizer_config.get('dense_optimizer', 'rmsprop')
        if kind != 'rmsprop':
            raise NotImplementedError(f"dense_optimizer={kind!r}: only 'rmsprop' (OT/config.py:40) has a kernel")
        oc = dict(self.config.optimizer_config)
        oc.setdefault('dense_lr', oc.get('learning_rate', self.config.learning_rate))
        opt = ClipRMSprop(self.grads, lr=oc['dense_lr'], rho=oc.get('rho', 0.9), momentum=oc.get('momentum', 0.0),
                          eps=oc.get('epsilon', 1e-7), clip_norm=getattr(self.config, 'gradient_clip_norm', 0.0))
        return opt

    def _to_device(self, batch):
        return tuple({k: v.to(self.device, non_blocking=True) for k, v in part.items()} for part in batch)

    def train_step(self, batch_data) -> Dict[str, torch.Tensor]:
        """OT/train.py:111-155: forward, summed BCE, backward, clip, RMSprop, metric update.  Returns device tensors."""
        non_seq, seq, labels = self._to_device(batch_data)
        self.grads.zero()
        loss, probs = self.model.forward_with_loss(non_seq, seq, labels, training=True)
        loss.backward()
        self.grads.all_reduce(self.world_size)
        self.optimizer.step()
        with torch.no_grad():
            self.train_metrics.update_state(labels, {t: probs[t].detach() for t in self.config.tasks})
            task_losses = {t: bce_loss({t: probs[t].detach()}, labels, [t]) for t in self.config.tasks}
        return {'total_loss': loss.detach(), 'task_losses': task_losses}

    @torch.no_grad()
    def val_step(self, batch_data) -> Dict[str, torch.Tensor]:
        """OT/train.py:157-184."""
        non_seq, seq, labels = self._to_device(batch_data)
        loss, probs = self.model.forward_with_loss(non_seq, seq, labels, training=False)
        self.val_metrics.update_state(labels, probs)
        return {'total_loss': loss, 'task_losses': {t: bce_loss({t: probs[t]}, labels, [t]) for t in self.config.tasks}}

    def train(self, train_loader, val_loader=None, epochs: int = 10, save_freq: int = 1, early_stopping_patience: int = 5,
              log_every: int = 100) -> Dict:
        """OT/train.py:186-279.  Losses stay on the device during an epoch (one read-back per epoch, not per step)."""
        import time
        best_val_loss, patience_counter = float('inf'), 0
        train_dataset = train_loader.get_train_dataset()
        val_dataset = val_loader.get_val_dataset() if val_loader else None
        for epoch in range(epochs):
            start = time.time()
            self.model.train()
            train_losses = []
            for step, batch in enumerate(train_dataset):
                info = self.train_step(batch)
                train_losses.append(info['total_loss'])
                if log_every and step % log_every == 0:
                    print(f"Epoch {epoch + 1}, Step {step}, Loss: {float(info['total_loss']):.4f}")
            avg_train_loss = float(torch.stack(train_losses).mean())
            if val_dataset is not None:
                self.model.eval()
                avg_val_loss = float(torch.stack([self.val_step(b)['total_loss'] for b in val_dataset]).mean())
            else:
                avg_val_loss = avg_train_loss                                    # :222-223
            self.history['train_loss'].append(avg_train_loss)
            self.history['val_loss'].append(avg_val_loss)
            self.history['train_metrics'][epoch] = self.train_metrics.result()   # :233-237
            self.history['val_metrics'][epoch] = self.val_metrics.result() if val_dataset is not None else {}
            print(f'Epoch {epoch + 1}/{epochs} - Train Loss: {avg_train_loss:.4f}, Val Loss: {avg_val_loss:.4f}, Time: {time.time() - start:.2f}s')
            if avg_val_loss < best_val_loss:                                     # :245-252
                best_val_loss, patience_counter = avg_val_loss, 0
                self.save_model('best_model')
            else:
                patience_counter += 1
            if (epoch + 1) % save_freq == 0:                                     # :254-256
                self.save_model(f'model_epoch_{epoch + 1}')
            if patience_counter >= early_stopping_patience:                      # :258-261
                print(f'early stopping at epoch {epoch + 1}')
                break
            self.train_metrics.reset_states()                                    # :263-267
            self.val_metrics.reset_states()
        self.save_model('final_model')                                           # :269-270
        return self.history

    def save_model(self, model_name: str) -> None:
        """OT/train.py:281-314: weights, ``config.json`` and ``training_history.json`` under ``model_dir / model_name``."""
        import json
        from . import state
        if self.world_size > 1 and dist.is_initialized() and dist.get_rank() != 0:
            return                                                               # data parallel: replicas are identical, rank 0 writes
        path = self.model_dir / model_name
        path.mkdir(parents=True, exist_ok=True)
        state.save_weights(self.model, path / state.WEIGHTS_FILE)
        with open(path / 'config.json', 'w') as f:
            json.dump(self.config.to_dict(), f, indent=2)
        with open(path / 'training_history.json', 'w') as f:
            json.dump(self.history, f, indent=2)

    def load_model(self, model_path) -> None:
        """OT/train.py:316-338: config, fresh model, weights, history - each only if its file is there."""
        import json
        from pathlib import Path
        from . import state
        from .config import OneTransConfig
        from .model import OneTransModel
        model_path = Path(model_path)
        if (model_path / 'config.json').exists():
            with open(model_path / 'config.json') as f:
                self.config = OneTransConfig.from_dict(json.load(f))
        self.model = OneTransModel(self.config).to(self.device)
        if (model_path / state.WEIGHTS_FILE).exists():
            state.load_weights(self.model, model_path / state.WEIGHTS_FILE)
        broadcast_parameters(self.model, self.world_size)
        self.grads = FlatGradBuffer(self.model.parameters())
        self.optimizer = self._create_optimizer()
        if (model_path / 'training_history.json').exists():
            with open(model_path / 'training_history.json') as f:
                self.history = json.load(f)


def train_one_trans_model(config_name: str = 'small', data_paths: Optional[Dict[str, str]] = None, epochs: int = 10,
                          model_dir: str = './models', config=None) -> OneTransTrainer:
    """OT/train.py:341-375: config -> loader (sample data when no paths are given) -> trainer -> ``train`` with the same loader
    for training and validation; returns the trainer."""
    from .config import get_model_config
    from .data import DataLoader
    config = config if config is not None else get_model_config(config_name)
    data_loader = DataLoader(config)
    if data_paths:
        data_loader.load_datasets(data_paths.get('train'), data_paths.get('val'), data_paths.get('test'))
    else:
        data_loader.train_dataset = data_loader.create_sample_data(seed=1)
        data_loader.val_dataset = data_loader.create_sample_data(seed=2)
    trainer = OneTransTrainer(config, model_dir)
    trainer.train(train_loader=data_loader, val_loader=data_loader, epochs=epochs)
    return trainer


def _cli(argv=None) -> int:
    """``python -m recommend_b200.train`` - the command line of OT/train.py:378-419 (same flags; ``--data_dir`` files are a stub in
    the reference too, so both paths train on generated sample data)."""
    import argparse
    ap = argparse.ArgumentParser(description='Train a OneTrans model on the sm_100a path')
    ap.add_argument('--config', type=str, default='small', help='model preset: small, base, large')
    ap.add_argument('--epochs', type=int, default=10)
    ap.add_argument('--batch_size', type=int, default=32)
    ap.add_argument('--model_dir', type=str, default='./models')
    ap.add_argument('--data_dir', type=str, default=None)
    args = ap.parse_args(argv)
    if not torch.cuda.is_available():
        print('recommend_b200.train needs a CUDA device (sm_100a); there is no CPU fallback')
        return 2
    from .config import get_model_config
    config = get_model_config(args.config)
    config.batch_size = args.batch_size                                                   # OT/train.py:404-405
    trainer = train_one_trans_model(args.config, None, args.epochs, args.model_dir, config=config)
    print(f'training finished; models under {args.model_dir}; final loss {trainer.history["train_loss"][-1]:.4f}')
    return 0


if __name__ == '__main__':
    raise SystemExit(_cli())
