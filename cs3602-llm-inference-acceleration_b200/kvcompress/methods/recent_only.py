"""Sliding-window control group (reference methods/recent_only.py:16-70)."""

from typing import List, Tuple

import torch

from .. import _planner
from ._common import as_layer_list, cached_plans, execute, seq_lens


def recent_only_compress(past_key_values, window_size: int = 512, skip_layers: List[int] = [0, 1],
                         **kwargs) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """Keep the last ``window_size`` tokens.  Like the reference this returns *views* of the input
    (``x[:, :, -window_size:, :]``), so it moves no bytes and launches nothing."""
    layers = as_layer_list(past_key_values)
    plans = cached_plans(_planner.plan_recent_only, seq_lens(layers), window_size, skip_layers=skip_layers)
    return execute(layers, plans)


__all__ = ["recent_only_compress"]
