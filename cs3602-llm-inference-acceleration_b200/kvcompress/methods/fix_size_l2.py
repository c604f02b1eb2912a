"""Fixed-budget L2 eviction on sm_100a (reference methods/fix_size_l2.py:15-154)."""

from typing import List, Tuple

import torch

from .. import _planner
from ._common import as_layer_list, cached_plans, execute, seq_lens, stored_norms


def _random_indices(keys: torch.Tensor, zone_end: int, count: int) -> torch.Tensor:
    """strategy="random": the reference draws one ``torch.randperm`` per (batch, head) from the
    device generator (fix_size_l2.py:118-124) and sorts the kept indices (:129).  The same calls in
    the same order are made here so the RNG stream — and therefore the kept set — is identical."""
    batch, heads = keys.size(0), keys.size(1)
    picked = torch.stack([
        torch.stack([torch.randperm(zone_end, device=keys.device)[:count] for _ in range(heads)])
        for _ in range(batch)
    ])
    picked, _ = torch.sort(picked, dim=-1)
    return picked.to(torch.int32).contiguous()


def fix_size_l2_compress(past_key_values, fix_kv_size: int = 1024, keep_ratio: float = 0.0,
                         strategy: str = "keep_low", skip_layers: List[int] = [0, 1],
                         **kwargs) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """Cap every layer at ``fix_kv_size`` tokens: the last ``int(fix_kv_size * keep_ratio)`` are
    protected, the rest of the budget is chosen from the older tokens by ``strategy``
    ("keep_low" / "keep_high" L2 norm, or "random").  Unknown strategies raise ``ValueError``."""
    layers = as_layer_list(past_key_values)
    plans = cached_plans(_planner.plan_fix_size, seq_lens(layers), fix_kv_size, keep_ratio, strategy,
                         skip_layers=skip_layers)
    given = None
    if strategy == "random":
        given = {li: _random_indices(layers[li][0], p.sel_hi, p.k_sel)
                 for li, p in enumerate(plans) if p.score == _planner.SCORE_GIVEN_INDEX}
    return execute(layers, plans, given_indices=given, norms=stored_norms(past_key_values),
                   non_blocking=kwargs.get("non_blocking", False), output_device=kwargs.get("output_device"))


__all__ = ["fix_size_l2_compress"]
