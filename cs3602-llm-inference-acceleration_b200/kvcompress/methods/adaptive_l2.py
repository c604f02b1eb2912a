"""Sequence-length-adaptive L2 compression (reference methods/adaptive_l2.py:20-201)."""

from typing import List, Tuple

import torch

from .. import _planner
from ._common import as_layer_list, cached_plans, execute, seq_lens, stored_norms


def adaptive_l2_compress(past_key_values, target_size: int = 512, soft_limit: int = 256, hard_limit: int = 1024,
                         keep_ratio_min: float = 0.3, keep_ratio_max: float = 0.9, skip_layers: List[int] = [],
                         **kwargs) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """``S <= soft_limit``: untouched; ``soft < S <= hard``: keep ratio interpolated between
    ``keep_ratio_max`` and ``keep_ratio_min`` with the last 20 % protected; ``S > hard_limit``:
    4 sinks + lowest-norm middle + ``target_size // 2`` recent tokens."""
    layers = as_layer_list(past_key_values)
    if not layers:
        return layers
    plans = cached_plans(_planner.plan_adaptive, seq_lens(layers), target_size, soft_limit, hard_limit, keep_ratio_min,
                         keep_ratio_max, skip_layers=skip_layers)
    return execute(layers, plans, norms=stored_norms(past_key_values),
                   non_blocking=kwargs.get("non_blocking", False), output_device=kwargs.get("output_device"))


__all__ = ["adaptive_l2_compress"]
