"""Shared entry sequence of the eight compress functions."""

from typing import List, Sequence, Tuple

import torch

from .. import _engine
from ..utils import normalize_kv_cache


def as_layer_list(past_key_values) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """Shallow list copy of the cache, exactly as every reference method starts
    (e.g. l2_compress.py:46): the caller's list is never mutated."""
    return list(normalize_kv_cache(past_key_values))


def seq_lens(layers: Sequence[Tuple[torch.Tensor, torch.Tensor]]) -> List[int]:
    return [layer[0].size(2) for layer in layers]


def execute(layers, plans, given_indices=None):
    return _engine.run_plans(layers, plans, given_indices=given_indices)
