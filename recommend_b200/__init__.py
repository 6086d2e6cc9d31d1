"""recommend_b200 — B200-native (sm_100a) implementation of the OneTrans ranking path of
ScottHCL/recommend (``rank/scaling_up/oneTrans/practice``).  Exports mirror the reference package
(``OT/__init__.py:9-26``) for the classes on the hot path."""
__version__ = "0.1.0"

from .config import OneTransConfig, OneTransSmallConfig, OneTransLargeConfig, get_model_config
from .schedule import PyramidScheduler, resolve_keep_lens
from .model import (RMSNorm, MixedMHA, MixedFFN, OneTransBlock, Tokenizer, OneTransModel, TaskHead,
                    create_onetrans_model)
from .state import load_reference_style_params, export_reference_style_params
from .embedding import EventEmbedding, SparseAdagrad
from .inference import OneTransInferenceEngine, score_candidates_sharded, shard_bounds, gather_shards
from .metrics import BinaryTaskMetrics, exact_auc, user_auc
from .evaluate import OneTransEvaluator, load_model_for_evaluation, evaluate_model
from .train import OneTransTrainer, train_one_trans_model
from .data import DataLoader, OneTransDataset, FeatureProcessor, SequenceProcessor, create_sample_batch

__all__ = [
    'OneTransModel', 'OneTransConfig', 'OneTransSmallConfig', 'OneTransLargeConfig', 'get_model_config',
    'RMSNorm', 'MixedMHA', 'MixedFFN', 'OneTransBlock', 'Tokenizer', 'PyramidScheduler', 'TaskHead',
    'create_onetrans_model', 'resolve_keep_lens', 'load_reference_style_params', 'export_reference_style_params',
    'EventEmbedding', 'SparseAdagrad', 'OneTransInferenceEngine', 'BinaryTaskMetrics', 'exact_auc', 'user_auc', 'OneTransEvaluator',
    'load_model_for_evaluation', 'evaluate_model', 'OneTransTrainer', 'train_one_trans_model', 'DataLoader', 'OneTransDataset',
    'FeatureProcessor', 'SequenceProcessor', 'create_sample_batch', 'score_candidates_sharded', 'shard_bounds', 'gather_shards',
]
