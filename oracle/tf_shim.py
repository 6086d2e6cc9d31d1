"""A stand-in for the ~30 TensorFlow / Keras entry points the reference's OT/model.py calls — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Why: the reference cannot be imported here (``tensorflow==2.12.0`` is neither vendored nor installable), so the oracle
was a line-by-line restatement pinned only by hand KATs.  With this module registered as ``sys.modules['tensorflow']`` the reference's
OWN, UNMODIFIED ``model.py`` / ``config.py`` import from /root/reference and run on torch CPU tensors: the control flow that decides
results - the per-position weight selection (OT/model.py:67-74, 84-92, 154-161), the concat orders (:90-92, 163, 235, 253, 277), the
mask (:59-65, 109-110), the [SEP] placement (:269-272), the pyramid indices and gathers (:287-302, 356-371), the last-token heads
(:384-391) - is then the reference's, executed, not restated.  What this file restates is only what each TensorFlow op computes
(documented semantics of TF 2.12): ``Dense`` = ``x @ kernel + bias`` with Glorot-uniform kernels and zero biases,
``activation='gelu'`` = exact erf form, ``Embedding`` = row lookup with U(-0.05, 0.05) rows, ``band_part(m, -1, 0)`` = lower triangle,
``tf.where`` / ``softmax`` / ``einsum`` / ``gather`` / ``concat`` / ``reshape`` as named, ``Dropout`` = inverted dropout when training.

Used by ``tests/golden/make_reference_golden.py`` (run in the build container, where /root/reference exists) to produce
``tests/golden/reference_golden.npz``; ``tests/test_reference_golden.py`` checks the oracle against those vectors everywhere.
Eager-mode behaviours that matter are kept: ``Layer.__call__`` runs ``build(input_shape)`` once with the first input's shape
(OT/model.py:59-60 builds the causal mask that way), ``tf.shape`` unpacks into Python integers, ``tf.concat`` refuses mixed dtypes and
``tf.gather`` refuses out-of-range indices (CPU kernels of TF raise InvalidArgumentError for both)."""
from __future__ import annotations

import math
import sys
import types
from typing import List, Optional, Sequence

import torch

FLOAT = torch.float64            # what tf.float32 maps to: double precision keeps the golden vectors free of rounding noise
_GEN = torch.Generator().manual_seed(0)


def set_seed(seed: int) -> None:
    _GEN.manual_seed(seed)


class InvalidArgumentError(Exception):
    pass


# ---- dtypes / tensors ---------------------------------------------------------------------------------------------------
float32 = FLOAT
int32 = torch.int64
Tensor = torch.Tensor


def _t(x, dtype=None):
    return x if isinstance(x, torch.Tensor) else torch.as_tensor(x, dtype=dtype or (FLOAT if isinstance(x, float) else None))


_VARIABLES = set()           # ids of tensors that are tf.Variables (weights); plain tensors such as a cached (k, v) are not


def _variable(t: torch.Tensor) -> torch.Tensor:
    _VARIABLES.add(id(t))
    return t


def Variable(initial_value, **_):
    return _variable(_t(initial_value).clone())


def ones(shape, dtype=None):
    return torch.ones(*[int(s) for s in shape], dtype=dtype or FLOAT)


def zeros(shape, dtype=None):
    return torch.zeros(*[int(s) for s in shape], dtype=dtype or FLOAT)


def cast(x, dtype):
    return _t(x).to(dtype)


def shape(x):
    return tuple(int(s) for s in x.shape)


def concat(values: Sequence[torch.Tensor], axis: int):
    kinds = {v.dtype for v in values}
    if len(kinds) > 1:
        raise InvalidArgumentError(f'ConcatOp: inputs of different dtypes {sorted(map(str, kinds))}')
    ranks = {v.dim() for v in values}
    if len(ranks) > 1:
        raise InvalidArgumentError(f'ConcatOp: ranks of all input tensors should match, got shapes {[tuple(v.shape) for v in values]}')
    return torch.cat(list(values), dim=axis)


def reshape(x, new_shape):
    return x.reshape(*[int(s) for s in new_shape])


def einsum(eq, *ops):
    return torch.einsum(eq, *ops)


def where(cond, x, y):
    x, y = _t(x, FLOAT), _t(y, FLOAT)
    return torch.where(cond, x.to(FLOAT), y.to(FLOAT))


def gather(params, indices, axis=0):
    idx = torch.as_tensor(list(indices) if not isinstance(indices, torch.Tensor) else indices, dtype=torch.long)
    n = params.shape[axis]
    if idx.numel() and (int(idx.min()) < 0 or int(idx.max()) >= n):
        raise InvalidArgumentError(f'GatherV2: indices[{int((idx >= n).nonzero()[0]) if (idx >= n).any() else 0}] = {int(idx.max())} is not in [0, {n})')
    return torch.index_select(params, axis, idx)


class _Math(types.SimpleNamespace):
    pass


math_ns = _Math(
    reduce_mean=lambda x, axis=None, keepdims=False: torch.mean(x, dim=axis, keepdim=keepdims),
    square=torch.square,
    rsqrt=torch.rsqrt,
    sqrt=lambda x: torch.sqrt(_t(x, FLOAT)),
)
linalg = types.SimpleNamespace(band_part=lambda m, lower, upper: _band_part(m, lower, upper))
nn = types.SimpleNamespace(softmax=lambda x, axis=-1: torch.softmax(x, dim=axis))
random = types.SimpleNamespace(
    uniform=lambda shape, minval=0, maxval=None, dtype=None: (
        torch.randint(int(minval), int(maxval), tuple(shape), generator=_GEN) if dtype is int32
        else torch.rand(*shape, generator=_GEN, dtype=FLOAT) * ((1.0 if maxval is None else maxval) - minval) + minval))


def _band_part(m, lower, upper):
    if lower != -1 or upper != 0:
        raise NotImplementedError('only band_part(m, -1, 0) (the lower triangle) is used by the reference')
    return torch.tril(m)


# ---- Keras ----------------------------------------------------------------------------------------------------------------
class Layer:
    """``tf.keras.layers.Layer``: ``__call__`` builds once from the first input's shape, then dispatches to ``call``."""

    def __init__(self, **_):
        self._built = False

    def build(self, input_shape):
        pass

    def __call__(self, *args, **kwargs):
        if not getattr(self, '_built', False):
            first = args[0] if args else next(iter(kwargs.values()))
            if isinstance(first, torch.Tensor):
                self.build(tuple(first.shape))
            self._built = True
        return self.call(*args, **kwargs)

    # what get_model_info (OT/model.py:399-408) walks
    @property
    def trainable_weights(self) -> List[torch.Tensor]:
        out: List[torch.Tensor] = []

        def visit(obj):
            if isinstance(obj, torch.Tensor):
                if id(obj) in _VARIABLES:
                    out.append(obj)
            elif isinstance(obj, Layer):
                for k, v in vars(obj).items():
                    if k not in ('causal_mask',):
                        visit(v)
            elif isinstance(obj, (list, tuple)):
                for v in obj:
                    visit(v)
            elif isinstance(obj, dict):
                for v in obj.values():
                    visit(v)
        for k, v in vars(self).items():
            if k not in ('causal_mask',):
                visit(v)
        return out


class Model(Layer):
    pass


def _gelu(x):
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))       # Keras 'gelu': approximate=False


_ACTIVATIONS = {None: lambda x: x, 'gelu': _gelu, 'sigmoid': torch.sigmoid}


class Dense(Layer):
    def __init__(self, units, activation=None, use_bias=True, **_):
        super().__init__()
        self.units, self.activation, self.use_bias = int(units), activation, use_bias
        self.kernel: Optional[torch.Tensor] = None
        self.bias: Optional[torch.Tensor] = None

    def build(self, input_shape):
        fan_in, fan_out = int(input_shape[-1]), self.units
        limit = math.sqrt(6.0 / (fan_in + fan_out))                  # glorot_uniform
        self.kernel = _variable((torch.rand(fan_in, fan_out, generator=_GEN, dtype=FLOAT) * 2.0 - 1.0) * limit)
        if self.use_bias:
            self.bias = _variable(torch.zeros(fan_out, dtype=FLOAT))

    def call(self, x):
        if x.dtype != FLOAT:
            raise InvalidArgumentError(f'MatMul: expected a float tensor, got {x.dtype}')
        y = x @ self.kernel
        if self.use_bias:
            y = y + self.bias
        return _ACTIVATIONS[self.activation](y)


class Reshape(Layer):
    def __init__(self, target_shape, **_):
        super().__init__()
        self.target_shape = [int(s) for s in target_shape]

    def call(self, x):
        return x.reshape(x.shape[0], *self.target_shape)


class Embedding(Layer):
    def __init__(self, input_dim, output_dim, **_):
        super().__init__()
        self.input_dim, self.output_dim = int(input_dim), int(output_dim)
        self.embeddings: Optional[torch.Tensor] = None

    def build(self, input_shape):
        self.embeddings = _variable(torch.rand(self.input_dim, self.output_dim, generator=_GEN, dtype=FLOAT) * 0.1 - 0.05)   # 'uniform'

    def call(self, ids):
        return self.embeddings[ids.long()]


class Dropout(Layer):
    def __init__(self, rate, **_):
        super().__init__()
        self.rate = float(rate)

    def call(self, x, training=False):
        if not training or self.rate <= 0.0:
            return x
        keep = (torch.rand(x.shape, generator=_GEN, dtype=torch.float32) >= self.rate).to(x.dtype)    # fp32 draws, as the oracle's _dropout
        return x * keep / (1.0 - self.rate)


class Sequential(Layer):
    def __init__(self, layers=None, **_):
        super().__init__()
        self.layers = list(layers or [])

    def call(self, x, training=False):
        for layer in self.layers:
            x = layer(x)
        return x


def one_hot(indices, depth):
    return torch.nn.functional.one_hot(torch.as_tensor(indices).long(), int(depth)).to(FLOAT)


def count_params(w) -> int:
    return int(w.numel())


def install() -> types.ModuleType:
    """Register this module as ``tensorflow`` (and the sub-module paths the reference spells out)."""
    me = sys.modules[__name__]
    tf = types.ModuleType('tensorflow')
    for name in ('float32', 'int32', 'Tensor', 'Variable', 'ones', 'zeros', 'cast', 'shape', 'concat', 'reshape', 'einsum', 'where', 'gather',
                 'linalg', 'nn', 'random', 'one_hot'):
        setattr(tf, name, getattr(me, name))
    tf.math = math_ns
    tf.errors = types.SimpleNamespace(InvalidArgumentError=InvalidArgumentError)
    tf.data = types.SimpleNamespace(Dataset=object, AUTOTUNE=-1)      # names OT/data_loader.py mentions in annotations at import time
    keras = types.ModuleType('tensorflow.keras')
    keras.layers = types.SimpleNamespace(Layer=Layer, Dense=Dense, Reshape=Reshape, Embedding=Embedding, Dropout=Dropout)
    keras.Model, keras.Sequential = Model, Sequential
    keras.backend = types.SimpleNamespace(count_params=count_params)
    tf.keras = keras
    sys.modules['tensorflow'] = tf
    sys.modules['tensorflow.keras'] = keras
    return tf
