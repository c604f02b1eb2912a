"""SnapKV-lite: pooled inverted-norm voting + observation window (reference methods/snapkv_lite.py:24-154)."""

from dataclasses import replace
from typing import List, Optional, Sequence, Tuple

import torch

from .. import _engine, _planner
from ._common import as_layer_list, cached_plans, execute, seq_lens, stored_norms


def snapkv_lite_compress(past_key_values, observation_window: int = 32, keep_size: int = 512,
                         pooling_kernel: int = 5, skip_layers: List[int] = [],
                         **kwargs) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """Keep the last ``observation_window`` tokens plus the ``keep_size - observation_window`` prefix
    tokens with the highest pooled score ``avg_pool1d(max_norm + 1e-6 - ||K||_2, pooling_kernel)``.

    ``obs_queries=`` (keyword-only extension read from ``**kwargs`` so the drop-in signature is unchanged; off by
    default — the reference has no queries): one ``[B, H*G, W, D]`` tensor per
    layer holding the query states of the last ``W = observation_window`` positions.  The prefix score then
    becomes the SnapKV vote ``softmax(q.K^T / sqrt(D)).sum(window queries, group heads)`` computed on the tensor
    cores, pooled, selected and gathered exactly as above — all in ONE launch (``_engine.snapkv_vote_compress``).
    ``obs_lse=`` (optional, with ``obs_queries``): per layer the ``[B, H*G, W]`` float32 log-sum-exp of those queries'
    attention rows, as returned by a flash-attention forward; the kernel then reads K once instead of twice."""
    obs_queries: Optional[Sequence[Optional[torch.Tensor]]] = kwargs.get("obs_queries")
    layers = as_layer_list(past_key_values)
    if not layers:
        return layers
    plans = cached_plans(_planner.plan_snapkv, seq_lens(layers), observation_window, keep_size, pooling_kernel,
                         skip_layers=skip_layers)
    if obs_queries is None:
        return execute(layers, plans, norms=stored_norms(past_key_values),
                   non_blocking=kwargs.get("non_blocking", False), output_device=kwargs.get("output_device"))
    if len(obs_queries) != len(layers):
        raise ValueError(f"obs_queries: {len(obs_queries)} entries for {len(layers)} layers")
    voted = [li for li, p in enumerate(plans) if p.kind == _planner.GATHER and p.k_sel > 0]
    for li in voted:
        if obs_queries[li] is None:
            raise ValueError(f"obs_queries[{li}] is missing for a layer that is compressed")
    plans = [replace(p, score=_planner.SCORE_GIVEN_SCORE) if li in voted else p for li, p in enumerate(plans)]
    return _engine.snapkv_vote_compress(layers, plans, obs_queries, observation_window, lse=kwargs.get("obs_lse"))


__all__ = ["snapkv_lite_compress"]
