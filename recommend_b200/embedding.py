"""ID front end of the behaviour-sequence tokenizer (north_star item 1).

The reference hands the tokenizer pre-embedded 64-d events (OT/model.py:217-219, 262-265; OT/data_loader.py:126-154 draws
them at random).  The paper builds them from ID embeddings (PAPER:89-109) and the repository's idiom for that is "several
Embedding lookups -> concat" (recall/bert_like/kuaiformer/practice/model.py:58-94).  ``EventEmbedding`` is that idiom in
front of ``OneTransModel``: ``ids [B, L, n_fields]`` -> ``events [B, L, n_fields * field_dim]`` (bf16), with sparse gradients
(scatter-add of the event gradients into an all-zero fp32 gradient table) and a sparse Adagrad update of the touched rows
(``sparse_optimizer: 'adagrad'``, ``sparse_lr: 0.1``, OT/config.py:39-47).  Everything runs in ``libonetrans_sm100.so``
(``ot_embed_gather_fwd / ot_embed_scatter_bwd / ot_embed_adagrad_step``); there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from . import _lib as L
from . import ops


def _params(mod: 'EventEmbedding', ids: torch.Tensor) -> L.EmbedParams:
    p = L.EmbedParams()
    p.table, p.field_off, p.field_rows = mod.table.data_ptr(), mod.field_off.data_ptr(), mod.field_rows.data_ptr()
    p.ids, p.n_events, p.n_fields, p.ef = ids.data_ptr(), ids.numel() // mod.n_fields, mod.n_fields, mod.field_dim
    return p


class _GatherFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, mod: 'EventEmbedding', ids: torch.Tensor):
        n_events = ids.numel() // mod.n_fields
        out = torch.empty(n_events, mod.n_fields * mod.field_dim, dtype=torch.bfloat16, device=ids.device)
        p = _params(mod, ids)
        p.events, p.ld_events, p.bad_ids = out.data_ptr(), out.stride(0), mod.bad_ids.data_ptr()
        ops._run('ot_embed_gather_fwd', L.load().ot_embed_gather_fwd, p, f'F{mod.n_fields}x{mod.field_dim}', 0.0,
                 n_events * mod.n_fields * (4.0 + 6.0 * mod.field_dim))
        ctx.mod, ctx.ids = mod, ids
        return out.view(*ids.shape[:-1], mod.n_fields * mod.field_dim)

    @staticmethod
    def backward(ctx, d_out):
        mod, ids = ctx.mod, ctx.ids
        if d_out is None:
            return None, None, None
        d2 = d_out.reshape(-1, mod.n_fields * mod.field_dim)
        if d2.dtype != torch.bfloat16 or d2.stride(-1) != 1:
            d2 = d2.to(torch.bfloat16).contiguous()
        p = _params(mod, ids)
        p.events, p.ld_events, p.grad = d2.data_ptr(), d2.stride(0), mod.grad_table.data_ptr()
        ops._run('ot_embed_scatter_bwd', L.load().ot_embed_scatter_bwd, p, f'F{mod.n_fields}x{mod.field_dim}', 0.0,
                 d2.shape[0] * mod.n_fields * (4.0 + 10.0 * mod.field_dim))
        mod._touched.append(ids)
        return None, None, None


class EventEmbedding(nn.Module):
    """``n_fields`` embedding tables of ``field_dim`` columns, looked up per event and concatenated:
    ``ids [..., n_fields]`` (int32) -> ``[..., n_fields * field_dim]`` bf16.  ``vocab_sizes[f]`` rows in table f.
    Rows are fp32 masters (Keras Embedding default init U(-0.05, 0.05)); the lookup rounds them to bf16."""

    def __init__(self, vocab_sizes: Sequence[int], field_dim: int = 16):
        super().__init__()
        if field_dim % 8:
            raise ValueError('field_dim must be a multiple of 8 (16-byte output chunks)')
        self.vocab_sizes = [int(v) for v in vocab_sizes]
        self.n_fields, self.field_dim = len(self.vocab_sizes), int(field_dim)
        total = sum(self.vocab_sizes)
        # a plain buffer, not a Parameter: its gradient is sparse and never goes through autograd / the flat dense buffer
        self.register_buffer('table', (torch.rand(total, field_dim) - 0.5) * 0.1)
        off = [0]
        for v in self.vocab_sizes[:-1]:
            off.append(off[-1] + v)
        self.register_buffer('field_off', torch.tensor(off, dtype=torch.int64))
        self.register_buffer('field_rows', torch.tensor(self.vocab_sizes, dtype=torch.int64))
        self.register_buffer('bad_ids', torch.zeros(1, dtype=torch.int32))
        self.register_buffer('_anchor', torch.zeros(()), persistent=False)
        self.grad_table: Optional[torch.Tensor] = None
        self._touched: List[torch.Tensor] = []

    def _ensure_grad(self) -> None:
        if self.grad_table is None or self.grad_table.device != self.table.device:
            self.grad_table = torch.zeros_like(self.table)

    def forward(self, ids: torch.Tensor) -> torch.Tensor:
        if not ids.is_cuda or not self.table.is_cuda:
            raise RuntimeError('EventEmbedding runs on CUDA tensors only (no CPU fallback)')
        if ids.shape[-1] != self.n_fields:
            raise ValueError(f'ids must end in {self.n_fields} fields, got {tuple(ids.shape)}')
        ids = ids.to(torch.int32).contiguous()
        self._ensure_grad()
        anchor = self._anchor.detach().requires_grad_(torch.is_grad_enabled())   # gives the lookup an autograd edge
        return _GatherFn.apply(anchor, self, ids)

    def out_of_vocabulary_count(self) -> int:
        """Number of ids outside their table seen so far (those lookups returned zero rows).  Synchronises."""
        return int(self.bad_ids.item())


class SparseAdagrad:
    """Keras ``Adagrad(learning_rate, initial_accumulator_value=0.1, epsilon=1e-7)`` restricted to the rows the last
    backward pass touched (identical to the dense rule: untouched rows have zero gradient and do not move)."""

    def __init__(self, emb: EventEmbedding, lr: float = 0.1, initial_accumulator_value: float = 0.1, eps: float = 1e-7):
        self.emb, self.lr, self.eps = emb, float(lr), float(eps)
        self.acc = torch.full_like(emb.table, float(initial_accumulator_value))
        self.stamp = torch.zeros(emb.table.shape[0], dtype=torch.int32, device=emb.table.device)
        self.step_id = 0

    @torch.no_grad()
    def step(self) -> None:
        emb = self.emb
        self.step_id = self.step_id % 0x7FFFFFF0 + 1
        for ids in emb._touched:
            p = _params(emb, ids)
            p.grad, p.acc, p.stamp = emb.grad_table.data_ptr(), self.acc.data_ptr(), self.stamp.data_ptr()
            p.step_id, p.lr, p.eps = self.step_id, self.lr, self.eps
            n = ids.numel()
            ops._run('ot_embed_adagrad_step', L.load().ot_embed_adagrad_step, p, f'F{emb.n_fields}x{emb.field_dim}', 0.0,
                     n * (4.0 + 24.0 * emb.field_dim))
        emb._touched.clear()
