"""The compress-function protocol — the reference's plugin boundary (methods/base.py:12-50)."""

from typing import List, Protocol, Tuple, runtime_checkable

import torch

from ..utils import normalize_kv_cache  # noqa: F401  (the reference re-defines it here)


@runtime_checkable
class CompressFn(Protocol):
    """``fn(past_key_values, **kwargs) -> List[Tuple[Tensor, Tensor]]``; tensors are [B, H, S, D]."""

    def __call__(self, past_key_values, **kwargs) -> List[Tuple[torch.Tensor, torch.Tensor]]:
        ...
