"""Serving wrapper with the method surface of the reference's ``OneTransInferenceEngine``
(OT/examples/inference_example.py:21-219): ``preprocess_input`` / ``single_inference`` / ``batch_inference`` /
``get_stats`` / ``reset_stats``, on top of the sm_100a model.  Additive: ``rank_candidates`` - the two-stage path of
north_star item 5 (PAPER:144-151): the user's sequence-side K/V are computed once and every candidate only runs its
NS tokens (``OneTransModel.build_kv_cache`` / ``score_candidates``).

Differences from the reference, all forced: weights come from ``model_weights.npz`` (Keras weight order, what
``OneTransTrainer.save_model`` writes) or a torch ``state_dict`` file ``model_weights.pt`` (the reference's ``.h5`` needs h5py /
Keras) next to the same ``config.json``; the model is called with two arguments (the
reference passes one tuple, which its own ``call`` signature rejects - SURVEY.md D7)."""
from __future__ import annotations

import json
import time
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple, Union

import torch

from .config import OneTransConfig
from .model import OneTransModel


def shard_bounds(n: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Rows ``[start, end)`` of rank ``rank`` when ``n`` candidates are split into ``world_size`` contiguous shards whose sizes
    differ by at most one (the first ``n % world_size`` ranks take the extra row)."""
    base, extra = divmod(n, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_shards(local: torch.Tensor, n: int, world_size: int, rank: int, group=None) -> torch.Tensor:
    """All-gather of per-rank score shards ``local [T, n_rank]`` (shard layout of ``shard_bounds``) into ``[T, n]`` on every rank:
    the one collective of sharded candidate scoring ("scores AllGather [8192]").  Shards are padded to the largest
    shard so that one ``all_gather`` moves them (NCCL over NVLink on GPUs; gloo in the CPU tests)."""
    import torch.distributed as dist
    if world_size == 1 or n == 0:
        return local
    T = local.shape[0]
    width = -(-n // world_size)
    send = local.new_zeros(T, width)
    send[:, :local.shape[1]] = local
    recv = local.new_empty(world_size, T, width)
    dist.all_gather(list(recv.unbind(0)), send, group=group)      # contiguous views of one buffer; works on NCCL and gloo alike
    parts = []
    for r in range(world_size):
        s, e = shard_bounds(n, world_size, r)
        parts.append(recv[r, :, :e - s])
    return torch.cat(parts, dim=1)


@torch.no_grad()
def score_candidates_sharded(model: OneTransModel, user_sequence_features: Dict[str, torch.Tensor], candidate_non_seq: Dict[str, torch.Tensor],
                             world_size: int = 1, rank: int = 0, group=None, return_logits: bool = False) -> Dict[str, torch.Tensor]:
    """BASELINE config 5 on several GPUs: candidates are the shard axis.  Every rank builds the user's per-layer
    K/V cache itself (1-2 MB of state, cheaper to recompute than to broadcast), scores rows ``shard_bounds(C, world, rank)`` of
    the candidate features and the probabilities are all-gathered, so every rank returns ``{task: [C, 1]}`` for all C."""
    C = next(iter(candidate_non_seq.values())).shape[0]
    s, e = shard_bounds(C, world_size, rank)
    model.build_kv_cache(user_sequence_features)
    tasks = list(model.config.tasks)
    if e > s:
        out = model.score_candidates({k: v[s:e] for k, v in candidate_non_seq.items()}, return_logits=return_logits)
        local = torch.stack([out[t].reshape(-1).float() for t in tasks])
    else:
        local = torch.zeros(len(tasks), 0, dtype=torch.float32, device=next(iter(candidate_non_seq.values())).device)
    full = gather_shards(local, C, world_size, rank, group)
    return {t: full[i].unsqueeze(1) for i, t in enumerate(tasks)}


class OneTransInferenceEngine:
    def __init__(self, model_or_path: Union[OneTransModel, str, Path], device: str = 'cuda', pad_sequences: bool = True):
        self.device = torch.device(device)
        self.pad_sequences = pad_sequences
        if isinstance(model_or_path, OneTransModel):
            self.model_path = None
            self.model = model_or_path.to(self.device).eval()
            self.config = self.model.config
        else:
            self.model_path = Path(model_or_path)
            self.model = None
            self.config = None
            self.load_model()
        self.reset_stats()

    # OT/examples/inference_example.py:38-60
    def load_model(self) -> None:
        config_path = self.model_path / 'config.json'
        if not config_path.exists():
            raise FileNotFoundError(f'config file not found: {config_path}')
        with open(config_path, 'r') as f:
            self.config = OneTransConfig.from_dict(json.load(f))
        from . import state
        npz_path, weights_path = self.model_path / state.WEIGHTS_FILE, self.model_path / 'model_weights.pt'
        if not npz_path.exists() and not weights_path.exists():
            raise FileNotFoundError(f'model weights not found: {npz_path} (written by OneTransTrainer.save_model) or {weights_path}')
        self.model = OneTransModel(self.config)
        if npz_path.exists():
            state.load_weights(self.model, npz_path)
        else:
            self.model.load_state_dict(torch.load(weights_path, map_location='cpu'))
        self.model = self.model.to(self.device).eval()

    # OT/examples/inference_example.py:62-92
    def preprocess_input(self, user_features: Dict, item_features: Dict, context_features: Dict,
                         sequence_features: Dict) -> Tuple[Dict, Dict]:
        """Merge the scalar feature groups; keep the last ``max_seq_len`` events of every sequence and (as the reference
        does) left-pad shorter ones with zero events up to ``max_seq_len`` unless ``pad_sequences=False``."""
        non_seq = {}
        for group in (user_features, item_features, context_features):
            non_seq.update(group)
        seq = {}
        L = self.config.max_seq_len
        for name, data in sequence_features.items():
            t = torch.as_tensor(data, dtype=torch.float32)
            if t.shape[0] > L:
                t = t[-L:]
            if self.pad_sequences and t.shape[0] < L:
                t = torch.cat([t.new_zeros(L - t.shape[0], t.shape[1]), t], dim=0)
            seq[name] = t
        return non_seq, seq

    def _to_batch(self, samples: Sequence[Tuple[Dict, Dict]]):
        non_seq = {k: torch.stack([torch.as_tensor(s[0][k], dtype=torch.float32).reshape(1) for s in samples]).to(self.device)
                   for k in samples[0][0]}
        seq = {k: torch.stack([s[1][k] for s in samples]).to(self.device, torch.bfloat16) for k in samples[0][1]}
        return non_seq, seq

    # OT/examples/inference_example.py:94-132
    def single_inference(self, user_features: Dict, item_features: Dict, context_features: Dict,
                         sequence_features: Dict) -> Dict[str, float]:
        return self.batch_inference([(user_features, item_features, context_features, sequence_features)])[0]

    # OT/examples/inference_example.py:134-184
    def batch_inference(self, batch_data: List[Tuple[Dict, Dict, Dict, Dict]]) -> List[Dict[str, float]]:
        start = time.time()
        try:
            non_seq, seq = self._to_batch([self.preprocess_input(*sample) for sample in batch_data])
            with torch.no_grad():
                preds = self.model(non_seq, seq, training=False)
            host = {task: p.float().cpu() for task, p in preds.items()}
            results = [{task: float(host[task][i, 0]) for task in host} for i in range(len(batch_data))]
            self._update_stats(True, (time.time() - start) * 1000.0 / len(batch_data), len(batch_data))
            return results
        except Exception:
            self._update_stats(False, 0.0, len(batch_data))
            raise

    # additive: two-stage scoring with the cross-candidate K/V cache
    def rank_candidates(self, user_sequence_features: Dict, candidate_non_seq: Dict[str, Sequence[float]]) -> Dict[str, List[float]]:
        """``user_sequence_features``: one user's behaviour sequences ``{name: [L_i, 64]}``; ``candidate_non_seq``: the eleven
        scalar features for C candidates ``{name: [C]}``.  Returns ``{task: [C probabilities]}``."""
        start = time.time()
        C = len(next(iter(candidate_non_seq.values())))
        try:
            _, seq = self.preprocess_input({}, {}, {}, user_sequence_features)
            seq = {k: v.unsqueeze(0).to(self.device, torch.bfloat16) for k, v in seq.items()}
            non_seq = {k: torch.as_tensor(v, dtype=torch.float32).reshape(C, 1).to(self.device) for k, v in candidate_non_seq.items()}
            with torch.no_grad():
                self.model.build_kv_cache(seq)
                preds = self.model.score_candidates(non_seq)
            out = {task: p.float().reshape(-1).cpu().tolist() for task, p in preds.items()}
            self._update_stats(True, (time.time() - start) * 1000.0 / C, C)
            return out
        except Exception:
            self._update_stats(False, 0.0, C)
            raise

    # OT/examples/inference_example.py:186-219
    def _update_stats(self, success: bool, latency: float, batch_size: int = 1) -> None:
        st = self.inference_stats
        st['total_requests'] += batch_size
        if success:
            st['successful_requests'] += batch_size
            st['avg_latency_ms'] = 0.1 * latency + 0.9 * st['avg_latency_ms']     # the reference's exponential average
        else:
            st['failed_requests'] += batch_size

    def get_stats(self) -> Dict:
        st = dict(self.inference_stats)
        st['success_rate'] = st['successful_requests'] / st['total_requests'] * 100 if st['successful_requests'] > 0 else 0.0
        return st

    def reset_stats(self) -> None:
        self.inference_stats = {'total_requests': 0, 'avg_latency_ms': 0.0, 'successful_requests': 0, 'failed_requests': 0}
