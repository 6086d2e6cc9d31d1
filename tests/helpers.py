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


# ---- cases at kernel-supported shapes whose EXPECTED outputs come from the reference's own code (tests/golden/make_reference_golden.py) ----
REFERENCE_KERNEL_CASES = {
    'G_d256_pyramid_on_1_block': dict(hidden_dim=256, num_heads=4, ffn_dim=256, num_layers=1, num_ns_tokens=2, pyramid_enabled=True,
                                      B=64, seq_lens=(10, 5, 7), seed=21),
    'H_d256_pyramid_off_2_blocks': dict(hidden_dim=256, num_heads=4, ffn_dim=256, num_layers=2, num_ns_tokens=2, pyramid_enabled=False,
                                        B=64, seq_lens=(9, 6, 4), seed=22),
}


# BASELINE config 1 (the reference's own CPU-runnable case) at full size; pyramid off because the reference cannot prune past one block (D2)
REFERENCE_CPU_CASES = {
    'C1_small_ns16_pyramid_off': dict(hidden_dim=256, num_heads=4, ffn_dim=1024, num_layers=6, num_ns_tokens=16, pyramid_enabled=False,
                                      B=32, seq_lens=(86, 84, 84), seed=31),
}


def reference_case_inputs(spec):
    """Weights and inputs of a REFERENCE_KERNEL_CASES entry, rebuilt from seeds (nothing but the reference's outputs is stored):
    oracle ``init_params`` + ``randomize_small_params`` with the GEMM weights rounded to bf16-representable values, ``synthetic_batch``
    with bf16-representable events.  ``head_literal`` alignment: the reference gives dedicated weights to positions < num_ns_tokens."""
    ocfg = O.OracleConfig(hidden_dim=spec['hidden_dim'], num_layers=spec['num_layers'], num_heads=spec['num_heads'], ffn_dim=spec['ffn_dim'],
                          num_ns_tokens=spec['num_ns_tokens'], ns_param_alignment='head_literal', pyramid_enabled=spec['pyramid_enabled'],
                          dropout_rate=0.0)
    P = O.init_params(ocfg, seed=spec['seed'])
    O.randomize_small_params(P, seed=spec['seed'] + 1)
    for k in P:
        if P[k].dim() >= 2 and 'ns_tokenizer' not in k and 'task_heads' not in k and 'sep_embedding' not in k:
            P[k] = P[k].to(torch.bfloat16).to(torch.float32)
    non_seq, seq, _ = O.synthetic_batch(ocfg, spec['B'], spec['seq_lens'], seed=1234 + spec['seed'])
    non_seq, seq = bf16_round_inputs(non_seq, seq)
    return ocfg, P, non_seq, seq


def reference_case_checksum(P, non_seq, seq) -> float:
    return float(sum(v.double().sum() for v in P.values()) + sum(v.double().sum() for v in non_seq.values()) + sum(v.double().sum() for v in seq.values()))
