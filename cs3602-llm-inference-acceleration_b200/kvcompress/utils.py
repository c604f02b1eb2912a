"""Cache-container helpers at the boundary of the compress path.

Same names and meaning as the reference's ``kvcompress/utils.py`` (:12-116).  The one
behavioural difference is deliberate: ``normalize_kv_cache`` also understands the
``DynamicCache`` of transformers >= 5 (no ``to_legacy_cache``; iteration yields
``(K, V, sliding_window)`` 3-tuples), which the reference's version cannot unpack.
"""

from typing import List, Tuple, Union

import torch

try:  # transformers is only needed for to_dynamic_cache
    from transformers import DynamicCache
except Exception:  # pragma: no cover - transformers is present in the target image
    DynamicCache = None


def normalize_kv_cache(past_key_values) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """Any supported cache format -> list of ``(K, V)`` pairs (reference utils.py:30-44)."""
    if hasattr(past_key_values, "to_legacy_cache"):
        return past_key_values.to_legacy_cache()
    items = list(past_key_values)
    if items and any(len(item) != 2 for item in items):
        items = [(item[0], item[1]) for item in items]
    return items


def _adopt_layers(past_key_values):
    """A ``DynamicCache`` whose layers hold the given tensors themselves.  ``cache.update`` on an empty layer
    concatenates with an empty tensor, i.e. copies every layer once more (transformers cache_utils.py:119-120);
    the compressed tensors are fresh already, so the layers can simply own them."""
    from transformers.cache_utils import DynamicLayer

    layers = []
    for keys, values in past_key_values:
        layer = DynamicLayer()
        if not hasattr(layer, "is_initialized") or not hasattr(layer, "keys"):
            return None  # a transformers version with a different layer object: use the portable path
        layer.dtype, layer.device = keys.dtype, keys.device
        layer.keys, layer.values = keys, values
        layer.is_initialized = True
        layers.append(layer)
    cache = DynamicCache()
    if getattr(cache, "layers", None) != []:
        return None
    cache.layers.extend(layers)
    return cache


def to_dynamic_cache(past_key_values: List[Tuple[torch.Tensor, torch.Tensor]]):
    """List of ``(K, V)`` pairs -> ``DynamicCache`` (reference utils.py:12-27, which goes through
    ``cache.update``).  The layers adopt the tensors without the extra copy when the installed transformers
    exposes its layer objects; otherwise the reference's ``cache.update`` path is used."""
    if DynamicCache is None:
        raise RuntimeError("transformers is required for to_dynamic_cache")
    try:
        cache = _adopt_layers(past_key_values)
        if cache is not None:
            return cache
    except Exception:
        pass
    cache = DynamicCache()
    for layer_idx, (keys, values) in enumerate(past_key_values):
        cache.update(keys, values, layer_idx)
    return cache


def get_cache_size_mb(past_key_values) -> float:
    """Total K+V bytes in MiB (reference utils.py:47-65)."""
    total = 0
    for keys, values in normalize_kv_cache(past_key_values):
        total += keys.element_size() * keys.nelement() + values.element_size() * values.nelement()
    return total / (1024 ** 2)


def get_cache_info(past_key_values) -> dict:
    """Layer count, per-layer lengths and size (reference utils.py:68-94)."""
    layers = normalize_kv_cache(past_key_values)
    if not layers:
        return {"num_layers": 0, "seq_lengths": [], "total_size_mb": 0}
    seq_lengths = [keys.size(2) for keys, _ in layers]
    return {
        "num_layers": len(layers),
        "seq_lengths": seq_lengths,
        "min_seq_len": min(seq_lengths),
        "max_seq_len": max(seq_lengths),
        "avg_seq_len": sum(seq_lengths) / len(seq_lengths),
        "total_size_mb": get_cache_size_mb(layers),
    }


def get_seq_len(past_key_values, layer_idx: int = 0) -> int:
    """Sequence length of one layer, 0 if absent (reference utils.py:97-116)."""
    layers = normalize_kv_cache(past_key_values)
    if not layers or layer_idx >= len(layers):
        return 0
    return layers[layer_idx][0].size(2)
