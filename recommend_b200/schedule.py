"""Pyramid schedule — ``PyramidScheduler`` of the reference (OT/model.py:280-302) plus the explicit
per-layer keep-length lists the build uses.  Pure host integer arithmetic, bit-exact
with the reference (``int(total * ratio)`` evaluated in Python double, KAT T1)."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence


class PyramidScheduler:
    """Same constructor and ``get_layer_config`` as the reference (OT/model.py:283-302)."""

    def __init__(self, config):
        self.config = config
        self.pyramid_ratios = config.pyramid_ratios

    def get_layer_config(self, layer_idx: int, total_seq_len: int) -> Dict:
        if not self.config.pyramid_enabled or layer_idx >= len(self.pyramid_ratios):
            return {'keep_ratio': 1.0, 'query_indices': None}                     # OT/model.py:289-290
        keep_ratio = self.pyramid_ratios[layer_idx]
        keep_len = max(1, int(int(total_seq_len) * keep_ratio))                    # OT/model.py:293 (+D5)
        query_indices = list(range(total_seq_len - keep_len, total_seq_len))       # OT/model.py:296
        return {'keep_ratio': keep_ratio, 'query_indices': query_indices, 'keep_len': keep_len}

    # ---- additive ----
    def keep_lens(self, total_seq_len: int) -> List[int]:
        return resolve_keep_lens(self.config, total_seq_len)


def keep_lens_reference_ratio(L0: int, num_layers: int, ratios: Sequence[float]) -> List[int]:
    """Reference ratios with repair D2: ``keep_l = max(1, int(L0 * r_l))`` of the ORIGINAL length, taken
    from the tail of the CURRENT sequence (``min(keep, cur)``); layers past the list keep everything."""
    out, cur = [], int(L0)
    for l in range(num_layers):
        if l < len(ratios):
            k = min(max(1, int(int(L0) * ratios[l])), cur)
        else:
            k = cur
        out.append(k)
        cur = k
    return out


def keep_lens_linear_to_ns(L0: int, num_layers: int, L_ns: int) -> List[int]:
    """``L_NS + ((N-1-l)(L0-L_NS))//N``: linear pruning down to the NS tokens (BASELINE.json config 2)."""
    return [L_ns + ((num_layers - 1 - l) * (L0 - L_ns)) // num_layers for l in range(num_layers)]


def keep_lens_halving(L0: int, num_layers: int, L_ns: int) -> List[int]:
    """Query set halved per block, floored at L_NS (BASELINE.json config 4)."""
    out, cur = [], int(L0)
    for _ in range(num_layers):
        cur = max(L_ns, cur // 2)
        out.append(cur)
    return out


def resolve_keep_lens(config, L0: int) -> List[int]:
    """Per-layer kept tail lengths for a layer-0 length ``L0``."""
    n = config.num_layers
    if not config.pyramid_enabled:
        return [int(L0)] * n
    explicit: Optional[List[int]] = getattr(config, 'pyramid_keep_lens', None)
    if explicit is not None:
        if len(explicit) != n:
            raise ValueError(f'pyramid_keep_lens has {len(explicit)} entries for {n} layers')
        raw = list(explicit)
    else:
        sched = getattr(config, 'pyramid_schedule', 'reference_ratio')
        if sched == 'reference_ratio':
            return keep_lens_reference_ratio(L0, n, config.pyramid_ratios)
        if sched == 'linear_to_ns':
            raw = keep_lens_linear_to_ns(L0, n, config.num_ns_tokens)
        elif sched == 'halving':
            raw = keep_lens_halving(L0, n, config.num_ns_tokens)
        else:
            raise ValueError(f'unknown pyramid_schedule {sched!r}')
    out, cur = [], int(L0)
    for k in raw:
        k = max(1, min(int(k), cur))
        out.append(k)
        cur = k
    return out
