"""Pyramid per-layer budgets with L2 selection (reference methods/pyramid_kv.py:26-185)."""

from typing import List, Tuple

import torch

from .. import _planner
from ._common import as_layer_list, cached_plans, execute, seq_lens, stored_norms


def pyramid_kv_compress(past_key_values, base_size: int = 512, layer_decay: float = 0.9, min_size: int = 64,
                        profile: str = "exponential", skip_layers: List[int] = [],
                        **kwargs) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """Layer ``i`` gets ``max(int(base_size * layer_decay**i), min_size)`` tokens ("exponential"; also
    "linear" / "constant"): ``min(4, t//8)`` sinks, the lowest-norm middle tokens and ``t//2`` recent ones."""
    layers = as_layer_list(past_key_values)
    if not layers:
        return layers
    plans = cached_plans(_planner.plan_pyramid, seq_lens(layers), base_size, layer_decay, min_size, profile,
                         skip_layers=skip_layers)
    return execute(layers, plans, norms=stored_norms(past_key_values),
                   non_blocking=kwargs.get("non_blocking", False), output_device=kwargs.get("output_device"))


__all__ = ["pyramid_kv_compress"]
