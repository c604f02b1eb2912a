"""StreamingLLM sinks + recent window on sm_100a (reference methods/streaming_llm.py:19-170)."""

from typing import List, Tuple

import torch

from .. import _planner
from ._common import as_layer_list, cached_plans, execute, seq_lens


def streaming_llm_compress(past_key_values, start_size: int = 4, recent_size: int = 508,
                           skip_layers: List[int] = [], **kwargs) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """Keep tokens ``[0, start_size)`` and the last ``recent_size`` tokens of every layer longer than
    ``start_size + recent_size``; one gather-compaction launch for the whole call.

    ``non_blocking=True`` (keyword extension of every compress function, host-resident caches only): return as soon as
    the launch is queued instead of synchronising; the pinned output tensors are complete once the current CUDA stream
    has been synchronised — the contract of ``tensor.to("cpu", non_blocking=True)``.
    ``output_device="cuda"`` (same scope): write the compressed cache to that GPU instead of pinned host memory — the
    kept rows of an offloaded cache cross PCIe once, host to device, and decoding continues on the GPU."""
    layers = as_layer_list(past_key_values)
    if not layers:
        return layers
    plans = cached_plans(_planner.plan_streaming, seq_lens(layers), start_size, recent_size, skip_layers=skip_layers)
    return execute(layers, plans, non_blocking=kwargs.get("non_blocking", False), output_device=kwargs.get("output_device"))


def evict_for_space(past_key_values, num_coming: int, start_size: int = 4, recent_size: int = 508,
                    skip_layers: List[int] = []) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """Make room for ``num_coming`` tokens before they arrive (reference streaming_llm.py:114-170)."""
    layers = as_layer_list(past_key_values)
    if not layers:
        return layers
    plans = cached_plans(_planner.plan_evict_for_space, seq_lens(layers), num_coming, start_size, recent_size,
                         skip_layers=skip_layers)
    return execute(layers, plans)


__all__ = ["streaming_llm_compress", "evict_for_space"]
