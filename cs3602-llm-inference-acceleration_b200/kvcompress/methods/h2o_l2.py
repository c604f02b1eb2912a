"""H2O with L2-norm heavy hitters on sm_100a (reference methods/h2o_l2.py:25-153)."""

from typing import List, Tuple

import torch

from .. import _planner
from ._common import as_layer_list, cached_plans, execute, seq_lens, stored_norms


def h2o_l2_compress(past_key_values, start_size: int = 4, heavy_hitter_size: int = 64, recent_size: int = 444,
                    skip_layers: List[int] = [], **kwargs) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """Keep ``start_size`` sink tokens, the ``heavy_hitter_size`` lowest-||K||_2 tokens of the middle
    and the last ``recent_size`` tokens of every layer longer than their sum."""
    layers = as_layer_list(past_key_values)
    if not layers:
        return layers
    plans = cached_plans(_planner.plan_h2o, seq_lens(layers), start_size, heavy_hitter_size, recent_size,
                         skip_layers=skip_layers)
    return execute(layers, plans, norms=stored_norms(past_key_values),
                   non_blocking=kwargs.get("non_blocking", False), output_device=kwargs.get("output_device"))


__all__ = ["h2o_l2_compress"]
