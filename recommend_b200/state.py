"""Parameter name map between the packed module parameters and reference-style names.

Reference-style dict (Keras ``[in, out]`` kernels; index 0 of a leading group dimension = shared weights,
1+j = dedicated weights of NS token j):
  tokenizer.ns_tokenizer.{kernel,bias}            OT/model.py:211-214
  tokenizer.seq_projections.{i}.{kernel,bias}     OT/model.py:217-219
  tokenizer.sep_embedding                         OT/model.py:222
  blocks.{l}.norm{1,2}.scale                      OT/model.py:174-175
  blocks.{l}.attention.{Wq,Wk,Wv} [G,d,d], .Wo    OT/model.py:38-57
  blocks.{l}.ffn.{W1,b1,W2,b2}                    OT/model.py:136-147
  output_norm.scale                               OT/model.py:322
  task_heads.{task}.{0,1}.{kernel,bias}           OT/model.py:325-330
"""
from __future__ import annotations

from typing import Dict

import torch


def _pairs(model):
    tok = model.tokenizer
    yield 'tokenizer.ns_tokenizer.kernel', tok.ns_kernel
    yield 'tokenizer.ns_tokenizer.bias', tok.ns_bias
    for i in range(len(tok.seq_kernels)):
        yield f'tokenizer.seq_projections.{i}.kernel', tok.seq_kernels[i]
        yield f'tokenizer.seq_projections.{i}.bias', tok.seq_biases[i]
    yield 'tokenizer.sep_embedding', tok.sep_embedding
    for l, blk in enumerate(model.blocks):
        b = f'blocks.{l}.'
        yield b + 'norm1.scale', blk.norm1.scale
        yield b + 'norm2.scale', blk.norm2.scale
        yield b + 'attention.Wo', blk.attention.Wo
        yield b + 'ffn.W1', blk.ffn.W1
        yield b + 'ffn.b1', blk.ffn.b1
        yield b + 'ffn.W2', blk.ffn.W2
        yield b + 'ffn.b2', blk.ffn.b2
    yield 'output_norm.scale', model.output_norm.scale
    for t, head in model.task_heads.items():
        yield f'task_heads.{t}.0.kernel', head.kernel0
        yield f'task_heads.{t}.0.bias', head.bias0
        yield f'task_heads.{t}.1.kernel', head.kernel1
        yield f'task_heads.{t}.1.bias', head.bias1


@torch.no_grad()
def load_reference_style_params(model, P: Dict[str, torch.Tensor]) -> None:
    """Copy a reference-style parameter dict into a OneTransModel (in place, keeps device/dtype)."""
    for name, p in _pairs(model):
        p.copy_(P[name].to(p.dtype).reshape(p.shape))
    for l, blk in enumerate(model.blocks):
        b = f'blocks.{l}.attention.'
        blk.attention.Wqkv.copy_(torch.cat([P[b + 'Wq'], P[b + 'Wk'], P[b + 'Wv']], dim=2).to(blk.attention.Wqkv.dtype))


@torch.no_grad()
def export_reference_style_params(model, grads: bool = False) -> Dict[str, torch.Tensor]:
    """Reference-style dict of the model's parameters (or of their gradients)."""
    def pick(p):
        t = p.grad if grads else p
        return None if t is None else t.detach().float().cpu().clone()
    out = {name: pick(p) for name, p in _pairs(model)}
    for l, blk in enumerate(model.blocks):
        b = f'blocks.{l}.attention.'
        t = pick(blk.attention.Wqkv)
        d = blk.attention.hidden_dim
        for i, n in enumerate(('Wq', 'Wk', 'Wv')):
            out[b + n] = None if t is None else t[:, :, i * d:(i + 1) * d].contiguous()
    return out
