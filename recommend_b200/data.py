"""Synthetic batches of the shapes the reference's loader produces (OT/data_loader.py:301-329 ``create_sample_batch``;
per-sample generators :126-154): 11 scalar non-sequence features ``[B, 1]``, behaviour sequences of pre-embedded events
``[B, L_i, 64]`` and Bernoulli(0.5) labels per task.  Seeded with a CPU generator so that every arm of ``bench.py`` (and
every rank, with ``seed = 1234 + rank``) sees reproducible inputs."""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import torch

from .config import OneTransConfig


def create_sample_batch(config=None, batch_size=None, seq_lens: Optional[Sequence[int]] = None, seed: int = 1234,
                        ns_mode: str = 'normal', dtype=torch.float32) -> Tuple[Dict[str, torch.Tensor], Dict[str, torch.Tensor], Dict[str, torch.Tensor]]:
    """``(non_seq_features, seq_features, labels)`` on the CPU.  ``ns_mode='ids'`` follows the reference literally (ids
    ``randint(0, 100)`` / ``randint(0, 1000)`` cast to float, context ``U[0, 1)``, OT/data_loader.py:309-316);
    ``'normal'`` draws N(0, 1) scalars (raw id magnitudes up to 1000 through a Dense make bf16 tolerances meaningless,
    SURVEY.md §8d).  Events are N(0, 1) (OT/data_loader.py:146)."""
    if config is None or isinstance(config, int):
        # the reference's call form ``create_sample_batch(batch_size=2, config=None)`` (OT/data_loader.py:301-329): ids as the
        # reference draws them (cast to float), one random length in [1, max_seq_len] per sequence (:320-321)
        config, batch_size = (batch_size if batch_size is not None else OneTransConfig()), (2 if config is None else config)
    if seq_lens is None:        # no explicit lengths = the reference's form, whichever way the two arguments were passed
        batch_size = 2 if batch_size is None else batch_size
        ns_mode = 'ids'
    g = torch.Generator().manual_seed(seed)
    fc = config.feature_config
    B = batch_size
    if seq_lens is None:
        seq_lens = [int(torch.randint(1, config.max_seq_len + 1, (), generator=g)) for _ in fc['sequence_features']]
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


class FeatureProcessor:
    """OT/data_loader.py:13-65: per-feature statistics from a table (``fit``), z-score + clip to [-3, 3] for numerical features
    (:47-57), one-hot for categorical ones (:59-65).  Host-side preparation (numpy / pandas in the reference); ``data`` is anything
    with ``columns`` / ``__getitem__`` returning array-likes (a pandas DataFrame, or a dict of arrays)."""

    NUMERICAL = ('price', 'age', 'ctr')                                                    # :39-41
    CATEGORICAL = ('user_id', 'item_id', 'category', 'brand', 'location', 'device')         # :43-45

    def __init__(self, config: OneTransConfig):
        self.config = config
        self.feature_stats: Dict[str, Dict[str, float]] = {}
        self.vocab_sizes: Dict[str, int] = {}

    def _get_numerical_features(self):
        return list(self.NUMERICAL)

    def _get_categorical_features(self):
        return list(self.CATEGORICAL)

    def fit(self, data) -> None:
        columns = list(data.columns) if hasattr(data, 'columns') else list(data.keys())
        for f in self.NUMERICAL:
            if f in columns:
                col = torch.as_tensor(list(data[f]), dtype=torch.float64)
                self.feature_stats[f] = {'mean': float(col.mean()), 'std': float(col.std(unbiased=True)) if col.numel() > 1 else float('nan'),
                                         'min': float(col.min()), 'max': float(col.max())}      # pandas .std() is the sample (n - 1) deviation
        for f in self.CATEGORICAL:
            if f in columns:
                self.vocab_sizes[f] = int(torch.as_tensor(list(data[f])).max() + 1)

    def process_numerical_feature(self, feature_name: str, values):
        v = torch.as_tensor(values, dtype=torch.float64)
        if feature_name not in self.feature_stats:
            return v                                                                           # :49-50
        st = self.feature_stats[feature_name]
        return ((v - st['mean']) / (st['std'] + 1e-8)).clamp(-3, 3)                          # :53-56

    def process_categorical_feature(self, feature_name: str, values):
        v = torch.as_tensor(values)
        if feature_name not in self.vocab_sizes:
            return v                                                                           # :61-62
        return torch.nn.functional.one_hot(v.long(), self.vocab_sizes[feature_name]).to(torch.float32)   # :64-65


class SequenceProcessor:
    """OT/data_loader.py:68-101: keep the most recent ``max_seq_len`` events, left-pad shorter sequences with zero events."""

    def __init__(self, config: OneTransConfig):
        self.config = config
        self.max_seq_len = config.max_seq_len

    def process_sequence(self, sequence_data: torch.Tensor, sequence_type: str = '') -> torch.Tensor:
        L, F = self.max_seq_len, self.config.seq_feature_dim
        t = torch.as_tensor(sequence_data, dtype=torch.float32)
        if t.numel() == 0:
            return torch.zeros(L, F)                                       # :76-77
        if t.shape[0] > L:
            return t[-L:]                                                  # :80-82
        return torch.cat([t.new_zeros(L - t.shape[0], t.shape[1]), t], 0)  # :84-89 (pad in front)

    def process_multi_sequences(self, sequences: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        return {k: self.process_sequence(v, k) for k, v in sequences.items()}


class OneTransDataset:
    """Synthetic stand-in with the surface of OT/data_loader.py:104-222 (``__len__``, ``__getitem__`` -> ``(non_seq, seq, labels)``,
    batched iteration): ``num_samples`` users with ragged behaviour sequences (length U{1..max_seq_len}, N(0,1) events, :141-149),
    all eleven scalar features (the reference's sample generator fills five, :130-136, fewer than its tokenizer reads) and
    Bernoulli(0.5) labels (:151-154).  ``batches()`` plays the role of ``get_tf_dataset`` (:195-222): pinned host tensors, sequences
    padded by ``SequenceProcessor``, events already in ``dtype`` (bf16 halves the host->device bytes)."""

    def __init__(self, config: OneTransConfig, data_path=None, num_samples: int = 1000, seed: int = 0, dtype=torch.bfloat16,
                 label_signal: float = 0.0):
        if data_path is not None:
            raise NotImplementedError('file loading is a stub in the reference too (OT/data_loader.py:119-123 creates sample data)')
        self.config, self.dtype = config, dtype
        self.sequence_processor = SequenceProcessor(config)
        g = torch.Generator().manual_seed(seed)
        n, fc = num_samples, config.feature_config
        self.non_seq_data = {name: torch.randn(n, generator=g) for name in config.ns_features}
        self.seq_lens = {name: torch.randint(1, config.max_seq_len + 1, (n,), generator=g) for name in fc['sequence_features']}
        self.seq_data = {name: torch.randn(n, config.max_seq_len, config.seq_feature_dim, generator=g) for name in fc['sequence_features']}
        for name, lens in self.seq_lens.items():       # zero the events in front of each user's own history (left padding)
            pad = torch.arange(config.max_seq_len).unsqueeze(0) < (config.max_seq_len - lens).unsqueeze(1)
            self.seq_data[name][pad] = 0.0
        # label_signal > 0 (additive): P(y = 1) = sigmoid(signal * (scalar feature + mean of the user's last event)) instead of the
        # reference's coin flips, so that loss / AUC can move in end-to-end tests
        first_seq = self.seq_data[fc['sequence_features'][0]]
        logit = label_signal * (self.non_seq_data[config.ns_features[0]] + first_seq[:, -1, :].mean(1) * 4.0)
        self.labels = {t: (torch.rand(n, generator=g) < torch.sigmoid(logit)).float() for t in config.tasks}

    def __len__(self) -> int:
        return next(iter(self.non_seq_data.values())).shape[0]

    def __getitem__(self, idx: int):
        L = self.config.max_seq_len
        seq = {k: self.sequence_processor.process_sequence(v[idx, L - int(self.seq_lens[k][idx]):]) for k, v in self.seq_data.items()}
        return ({k: v[idx] for k, v in self.non_seq_data.items()}, seq, {k: v[idx] for k, v in self.labels.items()})

    def batches(self, batch_size: int = 32, shuffle: bool = False, seed: int = 0, drop_last: bool = False):
        n = len(self)
        order = torch.randperm(n, generator=torch.Generator().manual_seed(seed)) if shuffle else torch.arange(n)
        pin = torch.cuda.is_available()
        for s in range(0, n, batch_size):
            idx = order[s:s + batch_size]
            if drop_last and idx.numel() < batch_size:
                break
            mk = (lambda t: t.pin_memory()) if pin else (lambda t: t)
            yield ({k: mk(v[idx].reshape(-1, 1).contiguous()) for k, v in self.non_seq_data.items()},
                   {k: mk(v[idx].to(self.dtype).contiguous()) for k, v in self.seq_data.items()},
                   {k: mk(v[idx].reshape(-1, 1).contiguous()) for k, v in self.labels.items()})


class _BatchIterable:
    def __init__(self, dataset: OneTransDataset, batch_size: int, shuffle: bool):
        self.dataset, self.batch_size, self.shuffle, self.epoch = dataset, batch_size, shuffle, 0

    def __iter__(self):
        self.epoch += 1
        return self.dataset.batches(self.batch_size, self.shuffle, seed=self.epoch)

    def __len__(self) -> int:
        return (len(self.dataset) + self.batch_size - 1) // self.batch_size


class DataLoader:
    """OT/data_loader.py:225-298: holder of the train / validation / test datasets; ``get_*_dataset(batch_size)`` return re-iterable
    batch streams of ``(non_seq_features, seq_features, labels)`` (train shuffled, a new order every epoch)."""

    def __init__(self, config: OneTransConfig):
        self.config = config
        self.train_dataset = self.val_dataset = self.test_dataset = None

    def load_datasets(self, train_path=None, val_path=None, test_path=None, num_samples: Sequence[int] = (1000, 1000, 1000)):
        self.train_dataset = OneTransDataset(self.config, train_path, num_samples[0], seed=1)
        self.val_dataset = OneTransDataset(self.config, val_path, num_samples[1], seed=2)
        self.test_dataset = OneTransDataset(self.config, test_path, num_samples[2], seed=3)

    def create_sample_data(self, num_samples: int = 1000, seed: int = 0, **kw) -> 'OneTransDataset':
        """What OT/train.py:360-362 and OT/evaluate.py:447-448 call on the loader (the reference class lacks it)."""
        return OneTransDataset(self.config, None, num_samples, seed, **kw)

    def _get(self, ds, what: str, batch_size, shuffle: bool):
        if ds is None:
            raise ValueError(f'{what} dataset not loaded')                  # :246-247, 257-258, 268-269
        return _BatchIterable(ds, batch_size or self.config.batch_size, shuffle)

    def get_train_dataset(self, batch_size: int = None):
        return self._get(self.train_dataset, 'train', batch_size, True)

    def get_val_dataset(self, batch_size: int = None):
        return self._get(self.val_dataset, 'validation', batch_size, False)

    def get_test_dataset(self, batch_size: int = None):
        return self._get(self.test_dataset, 'test', batch_size, False)

    def get_data_info(self) -> Dict:
        info = {}
        for key, ds in (('train_samples', self.train_dataset), ('val_samples', self.val_dataset), ('test_samples', self.test_dataset)):
            if ds:
                info[key] = len(ds)
        return info
