"""KnormPress ratio compression on sm_100a (reference methods/l2_compress.py:18-92)."""

from typing import List, Tuple

import torch

from .. import _planner
from ._common import as_layer_list, cached_plans, execute, seq_lens, stored_norms


def l2_compress(past_key_values, keep_ratio: float = 1.0, prune_after: int = 1000,
                skip_layers: List[int] = [0, 1], **kwargs) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """Keep the ``ceil(keep_ratio * S)`` lowest-||K||_2 tokens of every (batch, head), in temporal order.

    Same signature, defaults and results as the reference; layers with ``S <= prune_after``, layers
    in ``skip_layers`` and everything when ``keep_ratio >= 1`` are returned as the same tensor objects.
    """
    layers = as_layer_list(past_key_values)
    plans = cached_plans(_planner.plan_l2, seq_lens(layers), keep_ratio, prune_after, skip_layers=skip_layers)
    return execute(layers, plans, norms=stored_norms(past_key_values),
                   non_blocking=kwargs.get("non_blocking", False), output_device=kwargs.get("output_device"))


__all__ = ["l2_compress"]
