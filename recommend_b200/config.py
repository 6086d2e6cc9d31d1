"""OneTrans configuration — same class names, fields and defaults as the reference's
``rank/scaling_up/oneTrans/practice/config.py`` ("OT/config.py"), so a user of the reference finds the
same attribute bag.  Additive fields (never read by the reference) are grouped at the end."""
from __future__ import annotations

from typing import Dict, List, Optional


# Field -> (default, where the reference sets it).  One table instead of one assignment per line: the values are the
# reference's, the code is not.
_REFERENCE_DEFAULTS = (
    # model architecture: the paper's OneTrans-L
    ('hidden_dim', 384, 'OT/config.py:14'), ('num_layers', 8, ':15'), ('num_heads', 4, ':16'), ('ffn_dim', 1536, ':17'),
    # inputs
    ('max_seq_len', 2048, ':20'), ('num_ns_tokens', 12, ':21'), ('sep_token_id', 0, ':22'),
    # mixed parameterisation (dead flags in the reference, D21)
    ('shared_s_params', True, ':25'), ('dedicated_ns_params', True, ':26'),
    # pyramid stack
    ('pyramid_enabled', True, ':29'), ('pyramid_ratios', (0.5, 0.3, 0.2, 0.1, 0.05, 0.03, 0.02, 0.01), ':30'),
    # training
    ('batch_size', 2048, ':33'), ('learning_rate', 0.005, ':34'), ('num_epochs', 100, ':35'), ('warmup_steps', 10000, ':36'),
    # regularisation
    ('dropout_rate', 0.1, ':50'), ('weight_decay', 0.0, ':51'), ('gradient_clip_norm', 90.0, ':52'),
    # tasks
    ('tasks', ('ctr', 'cvr'), ':63'),
    # system flags (never read by OT/model.py, D20)
    ('use_mixed_precision', True, ':66'), ('use_kv_cache', True, ':67'), ('use_flash_attention', True, ':68'),
    ('use_activation_recompute', True, ':69'),
)
_OPTIMIZER_DEFAULTS = dict(dense_optimizer='rmsprop', sparse_optimizer='adagrad', dense_lr=0.005, sparse_lr=0.1, beta1=0.1, beta2=1.0,
                           momentum=0.99999)                                                  # OT/config.py:39-47
_FEATURE_GROUPS = dict(user_features=('user_id', 'age', 'gender', 'location'), item_features=('item_id', 'category', 'price', 'brand'),
                       context_features=('time', 'device', 'platform'),
                       sequence_features=('click_seq', 'cart_seq', 'purchase_seq'))       # OT/config.py:55-60
# additive fields: never read by the reference
_ADDITIVE_DEFAULTS = dict(
    seq_feature_dim=64,              # width of a pre-embedded event (OT/model.py:433-442)
    ns_param_alignment='tail',       # 'tail' (repair D4) | 'head_literal' (OT/model.py:69-74 as written)
    pyramid_keep_lens=None,          # explicit per-layer kept tail lengths
    pyramid_schedule='reference_ratio',   # 'reference_ratio' | 'linear_to_ns' | 'halving'
    ns_feature_names=None,           # features the NS Dense is built with (None = all)
    rms_eps=1e-6,                    # OT/model.py:14
    hp_ns_residual=True,             # fp32 residual stream for the NS-token rows
)


class OneTransConfig:
    """Mirror of ``OneTransConfig`` (OT/config.py:9-82): the same attribute bag with the same defaults."""

    _OVERRIDES: Dict[str, object] = {}

    def __init__(self):
        for name, default, _where in _REFERENCE_DEFAULTS:
            setattr(self, name, list(default) if isinstance(default, tuple) else default)
        self.optimizer_config = dict(_OPTIMIZER_DEFAULTS)
        self.feature_config = {k: list(v) for k, v in _FEATURE_GROUPS.items()}
        for name, default in _ADDITIVE_DEFAULTS.items():
            setattr(self, name, default)
        for klass in reversed(type(self).__mro__):          # size presets of the subclasses
            for name, value in getattr(klass, '_OVERRIDES', {}).items():
                setattr(self, name, value)

    # OT/config.py:71-82
    def to_dict(self) -> Dict:
        return {k: v for k, v in self.__dict__.items() if not k.startswith('_')}

    @classmethod
    def from_dict(cls, config_dict: Dict) -> 'OneTransConfig':
        config = cls()
        for key, value in config_dict.items():
            if hasattr(config, key):
                setattr(config, key, value)
        return config

    # helpers (additive)
    @property
    def ns_features(self) -> List[str]:
        if self.ns_feature_names is not None:
            return list(self.ns_feature_names)
        fc = self.feature_config
        return fc['user_features'] + fc['item_features'] + fc['context_features']   # OT/model.py:243-245

    @property
    def head_dim(self) -> int:
        return self.hidden_dim // self.num_heads                                     # OT/model.py:34


class OneTransSmallConfig(OneTransConfig):
    """OT/config.py:85-92 — the paper's OneTrans-S (d 256, 6 blocks, F 1024)."""

    _OVERRIDES = dict(hidden_dim=256, num_layers=6, ffn_dim=1024)


class OneTransLargeConfig(OneTransConfig):
    """OT/config.py:95-103 — NOT the paper's OneTrans-L."""

    _OVERRIDES = dict(hidden_dim=512, num_layers=12, num_heads=8, ffn_dim=2048)


def get_model_config(model_type: str = 'default') -> OneTransConfig:
    """OT/config.py:106-117.  ``'base'`` (the name the reference's CLI/README use, D8) aliases ``'default'``."""
    presets = dict(small=OneTransSmallConfig, default=OneTransConfig, base=OneTransConfig, large=OneTransLargeConfig)
    try:
        return presets[model_type]()
    except KeyError:
        raise ValueError(f'unknown model type {model_type!r}; expected one of {sorted(presets)}') from None


DEFAULT_CONFIG = OneTransConfig()
