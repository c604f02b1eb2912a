"""SnapKV-lite: pooled inverted-norm voting + observation window (reference methods/snapkv_lite.py:24-154)."""

from typing import List, Tuple

import torch

from .. import _planner
from ._common import as_layer_list, cached_plans, execute, seq_lens


def snapkv_lite_compress(past_key_values, observation_window: int = 32, keep_size: int = 512,
                         pooling_kernel: int = 5, skip_layers: List[int] = [],
                         **kwargs) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """Keep the last ``observation_window`` tokens plus the ``keep_size - observation_window`` prefix
    tokens with the highest pooled score ``avg_pool1d(max_norm + 1e-6 - ||K||_2, pooling_kernel)``."""
    layers = as_layer_list(past_key_values)
    if not layers:
        return layers
    plans = cached_plans(_planner.plan_snapkv, seq_lens(layers), observation_window, keep_size, pooling_kernel, skip_layers=skip_layers)
    return execute(layers, plans)


__all__ = ["snapkv_lite_compress"]
