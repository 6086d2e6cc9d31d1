"""Training-step host logic for the OneTrans path: BCE loss as the reference builds it
(OT/train.py:78-93, 124-128), a flat fp32 gradient buffer shared by every parameter, and the
data-parallel gradient all-reduce (PAPER:190; the reference code has no distributed strategy).

One process per GPU; ``torch.distributed`` (NCCL over NVLink/NVSwitch) is the plumbing."""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional

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

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view(p.shape)
            off += p.numel()

    def zero(self) -> None:
        self.flat.zero_()

    def all_reduce(self, world_size: int, bucket_bytes: int = 256 << 20) -> None:
        """Sum over ranks then average.  Buckets keep each NCCL call in its high-bandwidth regime."""
        if world_size <= 1:
            return
        n_per = max(1, bucket_bytes // 4)
        for s in range(0, self.flat.numel(), n_per):
            dist.all_reduce(self.flat[s:s + n_per], op=dist.ReduceOp.SUM)
        self.flat.mul_(1.0 / world_size)


def train_step(model, grads: FlatGradBuffer, non_seq, seq, labels, world_size: int = 1) -> torch.Tensor:
    """forward + BCE + backward (+ gradient all-reduce): the "fwd+bwd" of the headline metric
    (OT/train.py:116-131).  Returns the detached loss tensor (no host sync)."""
    grads.zero()
    probs = model(non_seq, seq, training=True)
    loss = bce_loss(probs, labels, model.config.tasks)
    loss.backward()
    grads.all_reduce(world_size)
    return loss.detach()
