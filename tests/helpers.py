"""Shared helpers for the parity tests: build an oracle (CPU fp32/fp64) and the CUDA model with the
same weights and inputs."""
import torch

from oracle import onetrans_oracle as O
import recommend_b200 as R


def make_configs(hidden_dim=256, num_layers=6, num_heads=4, ffn_dim=1024, num_ns_tokens=16, schedule='reference_ratio',
                 keep_lens=None, alignment='tail', pyramid_enabled=True):
    ocfg = O.OracleConfig(hidden_dim=hidden_dim, num_layers=num_layers, num_heads=num_heads, ffn_dim=ffn_dim,
                          num_ns_tokens=num_ns_tokens, ns_param_alignment=alignment, pyramid_enabled=pyramid_enabled)
    cfg = R.OneTransConfig()
    cfg.hidden_dim, cfg.num_layers, cfg.num_heads, cfg.ffn_dim = hidden_dim, num_layers, num_heads, ffn_dim
    cfg.num_ns_tokens, cfg.ns_param_alignment, cfg.pyramid_enabled = num_ns_tokens, alignment, pyramid_enabled
    cfg.dropout_rate = 0.0
    cfg.pyramid_schedule = schedule
    cfg.pyramid_keep_lens = keep_lens
    ocfg.dropout_rate = 0.0
    return ocfg, cfg


def oracle_keep_lens(ocfg, cfg, L0):
    """Give the oracle the same explicit schedule the model resolves."""
    ocfg.pyramid_keep_lens = R.resolve_keep_lens(cfg, L0)
    return ocfg.pyramid_keep_lens


def bf16_round_inputs(non_seq, seq):
    """The CUDA path reads behaviour events as bf16; round the oracle's copy the same way so the comparison
    measures the kernels, not the input cast."""
    return non_seq, {k: v.to(torch.bfloat16).to(v.dtype) for k, v in seq.items()}


def to_cuda(d, dtype=None):
    return {k: (v.to('cuda') if dtype is None else v.to('cuda', dtype)) for k, v in d.items()}


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / (max|b| + tiny)"""
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().cpu().flatten(), b.double().cpu().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))
