"""Synthetic batches of the shapes the reference's loader produces (OT/data_loader.py:301-329 ``create_sample_batch``;
per-sample generators :126-154): 11 scalar non-sequence features ``[B, 1]``, behaviour sequences of pre-embedded events
``[B, L_i, 64]`` and Bernoulli(0.5) labels per task.  Seeded with a CPU generator so that every arm of ``bench.py`` (and
every rank, with ``seed = 1234 + rank``) sees reproducible inputs."""
from __future__ import annotations

from typing import Dict, Sequence, Tuple

import torch

from .config import OneTransConfig


def create_sample_batch(config: OneTransConfig, batch_size: int, seq_lens: Sequence[int], seed: int = 1234,
                        ns_mode: str = 'normal', dtype=torch.float32) -> Tuple[Dict[str, torch.Tensor], Dict[str, torch.Tensor], Dict[str, torch.Tensor]]:
    """``(non_seq_features, seq_features, labels)`` on the CPU.  ``ns_mode='ids'`` follows the reference literally (ids
    ``randint(0, 100)`` / ``randint(0, 1000)`` cast to float, context ``U[0, 1)``, OT/data_loader.py:309-316);
    ``'normal'`` draws N(0, 1) scalars (raw id magnitudes up to 1000 through a Dense make bf16 tolerances meaningless,
    SURVEY.md §8d).  Events are N(0, 1) (OT/data_loader.py:146)."""
    g = torch.Generator().manual_seed(seed)
    fc = config.feature_config
    B = batch_size
    non_seq: Dict[str, torch.Tensor] = {}
    if ns_mode == 'ids':
        for n in fc['user_features']:
            non_seq[n] = torch.randint(0, 100, (B, 1), generator=g).to(dtype)
        for n in fc['item_features']:
            non_seq[n] = torch.randint(0, 1000, (B, 1), generator=g).to(dtype)
        for n in fc['context_features']:
            non_seq[n] = torch.rand(B, 1, generator=g, dtype=torch.float64).to(dtype)
    else:
        for n in config.ns_features:
            non_seq[n] = torch.randn(B, 1, generator=g, dtype=torch.float64).to(dtype)
    seq = {}
    for name, L in zip(fc['sequence_features'], seq_lens):
        seq[name] = torch.randn(B, L, config.seq_feature_dim, generator=g, dtype=torch.float64).to(dtype)
    labels = {t: (torch.rand(B, 1, generator=g) < 0.5).to(dtype) for t in config.tasks}
    return non_seq, seq, labels
