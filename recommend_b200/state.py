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


# ---- Keras weight order / checkpoint files (OT/train.py:281-338 save_model / load_model) -----------
def _keras_entries(model):
    """``(path, parameter, index)`` in the order of the reference model's ``model.weights`` / ``get_weights()``: Keras lists a
    layer's own variables, then its tracked sub-layers in attribute-assignment order - ``OneTransModel.__init__``
    (OT/model.py:308-333): tokenizer, blocks, output_norm, task_heads; ``Tokenizer`` (:206-222): ns_tokenizer, seq_projections,
    sep_embedding; ``OneTransBlock`` (:169-184): norm1, norm2, attention, ffn; ``MixedMHA`` (:29-57): Wq/Wk/Wv shared, then the
    Wq / Wk / Wv dedicated lists, Wo; ``MixedFFN`` (:128-147): ffn_shared, ffn_dedicated[j]; every Dense is kernel then bias.
    ``index`` selects a slice of a packed parameter: ``(group, part)`` for Wqkv (part 0/1/2 = q/k/v), ``(group,)`` for W1/b1/W2/b2.
    Dedicated weights j belong to NS token j (group 1 + j; SURVEY.md D4 'tail' alignment).  The order (names and shapes) is checked
    against the attribute order of a live reference model in tests/test_reference_golden.py."""
    tok = model.tokenizer
    yield 'tokenizer/ns_tokenizer/dense/kernel', tok.ns_kernel, None
    yield 'tokenizer/ns_tokenizer/dense/bias', tok.ns_bias, None
    for i in range(len(tok.seq_kernels)):
        yield f'tokenizer/seq_projections/{i}/kernel', tok.seq_kernels[i], None
        yield f'tokenizer/seq_projections/{i}/bias', tok.seq_biases[i], None
    yield 'tokenizer/sep_embedding/embeddings', tok.sep_embedding, 'sep'
    for l, blk in enumerate(model.blocks):
        b = f'blocks/{l}/'
        yield b + 'norm1/scale', blk.norm1.scale, None
        yield b + 'norm2/scale', blk.norm2.scale, None
        att, n_ns = blk.attention, blk.attention.Wqkv.shape[0] - 1
        for part, n in enumerate(('Wq', 'Wk', 'Wv')):
            yield b + f'attention/{n}_shared/kernel', att.Wqkv, (0, part)
        for part, n in enumerate(('Wq', 'Wk', 'Wv')):
            for j in range(n_ns):
                yield b + f'attention/{n}_dedicated/{j}/kernel', att.Wqkv, (1 + j, part)
        yield b + 'attention/Wo/kernel', att.Wo, None
        ffn = blk.ffn
        for g in range(ffn.W1.shape[0]):
            name = 'ffn_shared' if g == 0 else f'ffn_dedicated/{g - 1}'
            yield b + f'ffn/{name}/dense/kernel', ffn.W1, (g,)
            yield b + f'ffn/{name}/dense/bias', ffn.b1, (g,)
            yield b + f'ffn/{name}/dense_1/kernel', ffn.W2, (g,)
            yield b + f'ffn/{name}/dense_1/bias', ffn.b2, (g,)
    yield 'output_norm/scale', model.output_norm.scale, None
    # the reference keeps its heads in a plain dict attribute (OT/model.py:325-330); Keras' trackable dict wrapper flattens tracked
    # layers in sorted-key order, which for the default ['ctr', 'cvr'] coincides with insertion order
    for t in sorted(model.task_heads.keys()):
        head = model.task_heads[t]
        yield f'task_heads/{t}/dense/kernel', head.kernel0, None
        yield f'task_heads/{t}/dense/bias', head.bias0, None
        yield f'task_heads/{t}/dense_1/kernel', head.kernel1, None
        yield f'task_heads/{t}/dense_1/bias', head.bias1, None


def _keras_view(p: torch.Tensor, index) -> torch.Tensor:
    if index is None:
        return p
    if index == 'sep':                       # Embedding(1, d).embeddings is [1, d]
        return p.reshape(1, -1)
    if len(index) == 2:                      # Wqkv [G, d, 3d]
        d = p.shape[1]
        return p[index[0], :, index[1] * d:(index[1] + 1) * d]
    return p[index[0]]


@torch.no_grad()
def keras_weight_list(model):
    """``[(path, numpy array)]`` in the reference's ``model.get_weights()`` order, Keras shapes (``[in, out]`` kernels)."""
    return [(path, _keras_view(p.detach(), idx).float().cpu().numpy().copy()) for path, p, idx in _keras_entries(model)]


@torch.no_grad()
def load_keras_weight_list(model, arrays) -> None:
    """Inverse of ``keras_weight_list``: ``arrays`` is the list ``reference_model.get_weights()`` returns (or the arrays of an
    ``.h5`` weight file in topological order).  Shapes are checked one by one."""
    entries = list(_keras_entries(model))
    if len(arrays) != len(entries):
        raise ValueError(f'expected {len(entries)} weight arrays in Keras order, got {len(arrays)}')
    for (path, p, idx), a in zip(entries, arrays):
        view = _keras_view(p, idx)
        a = torch.as_tensor(a)
        if tuple(a.shape) != tuple(view.shape):
            raise ValueError(f'{path}: expected shape {tuple(view.shape)}, got {tuple(a.shape)}')
        view.copy_(a.to(view.dtype))
    for p in model.parameters():
        torch.autograd.graph.increment_version(p)      # bf16 compute copies refresh on the next forward


WEIGHTS_FILE = 'model_weights.npz'     # the reference writes model_weights.h5 (OT/train.py:286); h5py is not in this image


def save_weights(model, path) -> None:
    """All weights in Keras order into one ``.npz`` (keys ``'{position:04d}:{path}'``)."""
    import numpy as np
    np.savez(path, **{f'{i:04d}:{name}': a for i, (name, a) in enumerate(keras_weight_list(model))})


def load_weights(model, path) -> None:
    """Reads a file written by ``save_weights`` (names are checked against this model), or a TensorFlow-side
    ``np.savez(path, *reference_model.get_weights())`` (keys ``arr_0, arr_1, ...``: positional, shapes are checked one by one)."""
    import numpy as np
    with np.load(path) as z:
        if z.files and all(k.startswith('arr_') for k in z.files):
            load_keras_weight_list(model, [z[f'arr_{i}'] for i in range(len(z.files))])
            return
        keys = sorted(z.files, key=lambda k: int(k.split(':', 1)[0]))      # '<position>:<name>'; numeric, not lexicographic
        want = [name for name, _, _ in _keras_entries(model)]
        got = [k.split(':', 1)[1] for k in keys]
        if got != want:
            bad = next((i for i, (a, b) in enumerate(zip(got, want)) if a != b), min(len(got), len(want)))
            raise ValueError(f'{path}: weight names differ from this model at position {bad} '
                             f'({got[bad] if bad < len(got) else "<end>"} vs {want[bad] if bad < len(want) else "<end>"})')
        load_keras_weight_list(model, [z[k] for k in keys])
