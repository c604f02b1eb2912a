"""H2O with real attention scores (reference methods/h2o_attention.py:28-391).

SURVEY.md §8f lists this as a *next* row: it needs ``output_attentions=True`` model outputs, which
are outside the compress hot path.  It is provided so that every name in ``list_methods()``
resolves and behaves as in the reference:

* without a manager the reference falls back to exactly the ``h2o_l2`` selection
  (h2o_attention.py:337-351) — that runs on the fused sm_100a kernel;
* with a manager, the manager's head-summed top-k indices (shared by every batch entry and head,
  h2o_attention.py:194-213, :318-331) are handed to the gather kernel as caller-supplied indices.

The score bookkeeping itself (`H2OAttentionManager`) is small torch arithmetic on the model's
attention outputs and stays in torch.
"""

from typing import Dict, List, Optional, Tuple

import torch

from .. import _planner
from ._common import as_layer_list, cached_plans, execute, seq_lens


class H2OAttentionManager:
    """Decayed running sum of attention mass per key position (reference h2o_attention.py:28-213)."""

    def __init__(self, start_size: int = 4, heavy_hitter_size: int = 64, recent_size: int = 444,
                 num_layers: int = 32, num_heads: int = 32, decay_factor: float = 0.9, device=None):
        self.start_size = start_size
        self.heavy_hitter_size = heavy_hitter_size
        self.recent_size = recent_size
        self.total_cache_size = start_size + heavy_hitter_size + recent_size
        self.num_layers = num_layers
        self.num_heads = num_heads
        self.decay_factor = decay_factor
        self.device = device
        self.accumulated_attention: Dict[int, torch.Tensor] = {}
        self.token_positions: Dict[int, torch.Tensor] = {}
        self.current_seq_len = 0

    def reset(self) -> None:
        self.accumulated_attention = {}
        self.token_positions = {}
        self.current_seq_len = 0

    def update_attention_scores(self, attentions, skip_layers: List[int] = []) -> None:
        """acc <- decay * acc (zero-extended, or reset if the cache shrank) + attn.sum(query dim)  (:78-151)."""
        if attentions is None:
            return
        for layer_idx, attn in enumerate(attentions):
            if layer_idx in skip_layers or attn is None:
                continue
            batch, heads, _, key_len = attn.shape
            mass = attn.sum(dim=2)
            prev = self.accumulated_attention.get(layer_idx)
            if prev is None or prev.size(-1) > key_len:
                acc = torch.zeros(batch, heads, key_len, device=attn.device, dtype=attn.dtype)
            else:
                acc = prev * self.decay_factor
                if prev.size(-1) < key_len:
                    grow = torch.zeros(batch, heads, key_len - prev.size(-1), device=attn.device, dtype=attn.dtype)
                    acc = torch.cat([acc, grow], dim=-1)
            self.accumulated_attention[layer_idx] = acc + mass
            self.current_seq_len = key_len

    def get_heavy_hitter_indices(self, layer_idx: int, seq_len: int) -> torch.Tensor:
        """Indices (relative to the middle region, ascending) of its heaviest tokens  (:153-213)."""
        acc = self.accumulated_attention.get(layer_idx)
        if acc is None:  # no data: evenly spaced  (:167-177)
            middle_len = (seq_len - self.recent_size) - self.start_size
            if middle_len <= 0:
                return torch.tensor([], dtype=torch.long)
            step = max(1, middle_len // self.heavy_hitter_size)
            return torch.arange(0, middle_len, step)[: self.heavy_hitter_size]
        middle_end = min(seq_len, acc.size(-1)) - self.recent_size
        if middle_end <= self.start_size:
            return torch.tensor([], dtype=torch.long, device=acc.device)
        per_token = acc[:, :, self.start_size:middle_end].sum(dim=1)
        if acc.size(0) == 1:
            per_token = per_token.squeeze(0)
        k = min(self.heavy_hitter_size, middle_end - self.start_size)
        _, top = torch.topk(per_token, k, dim=-1)
        top, _ = torch.sort(top, dim=-1)
        return top


def h2o_attention_compress(past_key_values, attention_scores=None, h2o_manager: Optional[H2OAttentionManager] = None,
                           start_size: int = 4, heavy_hitter_size: int = 64, recent_size: int = 444,
                           skip_layers: List[int] = [], **kwargs) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """Sinks + heavy hitters + recent window; heavy hitters come from ``h2o_manager`` when given,
    otherwise from the lowest key norms (identical to ``h2o_l2_compress``)."""
    layers = as_layer_list(past_key_values)
    if not layers:
        return layers
    if h2o_manager is not None and attention_scores is not None:
        h2o_manager.update_attention_scores(attention_scores, skip_layers)  # :274-275
    plans = cached_plans(_planner.plan_h2o, seq_lens(layers), start_size, heavy_hitter_size, recent_size,
                         skip_layers=skip_layers)
    if h2o_manager is None:
        return execute(layers, plans)

    plans, given = manager_plans(layers, plans, h2o_manager, heavy_hitter_size)
    return execute(layers, plans, given_indices=given)


def manager_plans(layers, plans, h2o_manager: H2OAttentionManager, heavy_hitter_size: int):
    """Rewrite the h2o plans with the manager's heavy hitters (reference h2o_attention.py:318-331): per compressed
    layer the head-summed top-k rows become caller-supplied rows of the gather (``SCORE_GIVEN_INDEX``).  Returns
    (plans, {layer: int32 [B, H, count] absolute rows, ascending}).  Shared by the function and by
    ``KVSlabCache.compress_("h2o_attention", h2o_manager=...)``."""
    plans = list(plans)  # the manager rewrites per-layer plans below: never touch the cached set
    given = {}
    for li, plan in enumerate(plans):
        if plan.kind != _planner.GATHER or plan.region == 0:
            continue
        keys = layers[li][0]
        batch, heads, seq_len, _ = keys.shape
        middle_len = plan.sel_hi - plan.sel_lo
        picked = h2o_manager.get_heavy_hitter_indices(li, seq_len)
        count = min(len(picked), heavy_hitter_size, middle_len)  # :321
        if count > 0 and picked.dim() != 1:
            raise ValueError("h2o_attention: the manager's indices are shared across the batch; batch > 1 "
                             "is not supported (the reference fails on it as well)")
        rows = picked[:count].clamp(0, middle_len - 1).to(keys.device) + plan.sel_lo  # :324-326
        plans[li] = _planner.LayerPlan(_planner.GATHER, seq_len, sink=plan.sink, sel_lo=plan.sel_lo,
                                       sel_hi=plan.sel_hi, k_sel=count, tail=plan.tail,
                                       score=_planner.SCORE_GIVEN_INDEX if count > 0 else _planner.SCORE_NONE)
        if count > 0:
            given[li] = rows.to(torch.int32).view(1, 1, count).expand(batch, heads, count).contiguous()
    return plans, given


def create_h2o_manager_from_model(model, **kwargs) -> H2OAttentionManager:
    """Manager sized from ``model.config`` (reference h2o_attention.py:366-391)."""
    config = model.config
    return H2OAttentionManager(
        start_size=kwargs.get("start_size", 4),
        heavy_hitter_size=kwargs.get("heavy_hitter_size", 64),
        recent_size=kwargs.get("recent_size", 444),
        num_layers=getattr(config, "num_hidden_layers", 32),
        num_heads=getattr(config, "num_attention_heads", 32),
        decay_factor=kwargs.get("decay_factor", 0.9),
        device=next(model.parameters()).device,
    )


__all__ = ["H2OAttentionManager", "h2o_attention_compress", "create_h2o_manager_from_model", "manager_plans"]
