"""OneTrans configuration — same class names, fields and defaults as the reference's
``rank/scaling_up/oneTrans/practice/config.py`` ("OT/config.py"), so a user of the reference finds the
same attribute bag.  Additive fields (never read by the reference) are grouped at the end."""
from __future__ import annotations

from typing import Dict, List, Optional


class OneTransConfig:
    """Mirror of ``OneTransConfig`` (OT/config.py:9-82).  Defaults = the paper's OneTrans-L."""

    def __init__(self):
        # model architecture (OT/config.py:14-17)
        self.hidden_dim = 384
        self.num_layers = 8
        self.num_heads = 4
        self.ffn_dim = 1536
        # inputs (OT/config.py:20-22)
        self.max_seq_len = 2048
        self.num_ns_tokens = 12
        self.sep_token_id = 0
        # mixed parameterisation (OT/config.py:25-26; dead flags in the reference, D21)
        self.shared_s_params = True
        self.dedicated_ns_params = True
        # pyramid stack (OT/config.py:29-30)
        self.pyramid_enabled = True
        self.pyramid_ratios = [0.5, 0.3, 0.2, 0.1, 0.05, 0.03, 0.02, 0.01]
        # training (OT/config.py:33-36)
        self.batch_size = 2048
        self.learning_rate = 0.005
        self.num_epochs = 100
        self.warmup_steps = 10000
        # optimizer (OT/config.py:39-47)
        self.optimizer_config = {
            'dense_optimizer': 'rmsprop',
            'sparse_optimizer': 'adagrad',
            'dense_lr': 0.005,
            'sparse_lr': 0.1,
            'beta1': 0.1,
            'beta2': 1.0,
            'momentum': 0.99999,
        }
        # regularisation (OT/config.py:50-52)
        self.dropout_rate = 0.1
        self.weight_decay = 0.0
        self.gradient_clip_norm = 90.0
        # features (OT/config.py:55-60)
        self.feature_config = {
            'user_features': ['user_id', 'age', 'gender', 'location'],
            'item_features': ['item_id', 'category', 'price', 'brand'],
            'context_features': ['time', 'device', 'platform'],
            'sequence_features': ['click_seq', 'cart_seq', 'purchase_seq'],
        }
        # tasks (OT/config.py:63)
        self.tasks = ['ctr', 'cvr']
        # system flags (OT/config.py:66-69; never read by OT/model.py, D20)
        self.use_mixed_precision = True
        self.use_kv_cache = True
        self.use_flash_attention = True
        self.use_activation_recompute = True

        # ---- additive ----
        self.seq_feature_dim = 64                      # width of a pre-embedded event (OT/model.py:433-442)
        self.ns_param_alignment = 'tail'               # 'tail' (repair D4) | 'head_literal' (OT/model.py:69-74 as written)
        self.pyramid_keep_lens: Optional[List[int]] = None   # explicit per-layer kept tail lengths
        self.pyramid_schedule = 'reference_ratio'      # 'reference_ratio' | 'linear_to_ns' | 'halving'
        self.ns_feature_names: Optional[List[str]] = None    # features the NS Dense is built with (None = all)
        self.rms_eps = 1e-6                            # OT/model.py:14
        self.hp_ns_residual = True                     # fp32 residual stream for the NS-token rows

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

    def __init__(self):
        super().__init__()
        self.hidden_dim = 256
        self.num_layers = 6
        self.ffn_dim = 1024


class OneTransLargeConfig(OneTransConfig):
    """OT/config.py:95-103 — NOT the paper's OneTrans-L."""

    def __init__(self):
        super().__init__()
        self.hidden_dim = 512
        self.num_layers = 12
        self.num_heads = 8
        self.ffn_dim = 2048


def get_model_config(model_type: str = 'default') -> OneTransConfig:
    """OT/config.py:106-117.  ``'base'`` (the name the reference's CLI/README use, D8) aliases ``'default'``."""
    config_map = {
        'small': OneTransSmallConfig,
        'default': OneTransConfig,
        'base': OneTransConfig,
        'large': OneTransLargeConfig,
    }
    if model_type not in config_map:
        raise ValueError(f"未知的模型类型: {model_type}")
    return config_map[model_type]()


DEFAULT_CONFIG = OneTransConfig()
