"""Shared entry sequence of the eight compress functions."""

from typing import Callable, List, Sequence, Tuple

import torch

from .. import _engine
from ..utils import normalize_kv_cache

_PLAN_CACHE = {}
_PLAN_CACHE_MAX = 512


def as_layer_list(past_key_values) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """Shallow list copy of the cache, exactly as every reference method starts
    (e.g. l2_compress.py:46): the caller's list is never mutated."""
    return list(normalize_kv_cache(past_key_values))


def seq_lens(layers: Sequence[Tuple[torch.Tensor, torch.Tensor]]) -> List[int]:
    fast = _engine.fast_binding()
    if fast is not None and type(layers) is list:
        try:
            return fast.seq_lens(layers)
        except (TypeError, IndexError):
            pass
    return [layer[0].size(2) for layer in layers]


def cached_plans(plan_fn: Callable, lens: Sequence[int], *args, skip_layers=()) -> "_engine.PlanSet":
    """``plan_fn(lens, *args, skip_layers)`` as a :class:`_engine.PlanSet`, memoised on its arguments.
    A decode loop asks for the same plan every step (S = cap + 1 for every layer), so the planner's
    integer arithmetic and the packing of the launch records run once, not per step.  Plans are pure
    functions of these arguments; a planner error (``ValueError``) is never cached."""
    try:
        key = (plan_fn.__name__, tuple(lens), args, tuple(skip_layers))
        hit = _PLAN_CACHE.get(key)
    except TypeError:  # unhashable argument: plan without the cache
        return _engine.PlanSet(plan_fn(lens, *args, skip_layers))
    if hit is None:
        hit = _engine.PlanSet(plan_fn(lens, *args, skip_layers))
        if len(_PLAN_CACHE) >= _PLAN_CACHE_MAX:
            _PLAN_CACHE.clear()
        _PLAN_CACHE[key] = hit
    return hit


def stored_norms(past_key_values):
    """Per-layer key norms kept by the cache container (``KVSlabCache.key_norm_layers``), or ``None`` for plain
    ``(K, V)`` lists and HF caches: with them a method ranks rows without reading K (reference e.g.
    fix_size_l2.py:104-113 recomputes ``torch.norm`` on every call)."""
    fn = getattr(past_key_values, "key_norm_layers", None)
    return fn() if callable(fn) else None


def execute(layers, plans, given_indices=None, norms=None, non_blocking=False, output_device=None):
    return _engine.run_plans(layers, plans, given_indices=given_indices, norms=norms, non_blocking=non_blocking,
                             output_device=output_device)
