"""Token-by-token evaluation with attention-score H2O eviction — the caller loop of ``h2o_attention_compress``.

Same entry points, arguments and result keys as the reference's ``kvcompress/evaluate_attention.py:24-330``: every
forward pass runs with ``output_attentions=True``, the :class:`H2OAttentionManager` accumulates the attention mass
each cached position receives, and once the cache exceeds ``start_size + heavy_hitter_size + recent_size`` rows it is
compacted to sinks + heavy hitters + recent rows (the gather runs on the sm_100a kernels with the manager's rows as
``KVC_SCORE_GIVEN_INDEX``), after which the manager starts over — positions have moved (reference :164-177).

The model must be able to return attention weights (``attn_implementation="eager"``); ``tokenizer`` only needs
``encode(text, return_tensors="pt")`` — pass ``input_ids=`` to skip it (offline runs).
"""

from __future__ import annotations

import math
import time
from typing import Dict, List, Optional

import torch
from torch.nn import CrossEntropyLoss

from .evaluate import _EMPTY, _final_cache_size, evaluate_with_compression
from .methods.h2o_attention import H2OAttentionManager, h2o_attention_compress
from .utils import normalize_kv_cache, to_dynamic_cache


def evaluate_with_attention_compression(model, tokenizer=None, text: str = "",
                                        h2o_manager: Optional[H2OAttentionManager] = None, start_size: int = 4,
                                        heavy_hitter_size: int = 64, recent_size: int = 444, max_tokens: int = 3000,
                                        skip_layers: List[int] = [0, 1], device: Optional[torch.device] = None,
                                        show_progress: bool = True, *, input_ids: Optional[torch.Tensor] = None,
                                        return_nlls: bool = False) -> Dict[str, float]:
    """PPL / accuracy / TTFT / TPOT with attention-score H2O applied after every token."""
    if device is None:
        device = next(model.parameters()).device
    if h2o_manager is None:
        h2o_manager = H2OAttentionManager(start_size=start_size, heavy_hitter_size=heavy_hitter_size,
                                          recent_size=recent_size,
                                          num_layers=getattr(model.config, "num_hidden_layers", 32),
                                          num_heads=getattr(model.config, "num_attention_heads", 32), device=device)
    else:
        h2o_manager.reset()
    if input_ids is None:
        input_ids = tokenizer.encode(text, return_tensors="pt")
    input_ids = input_ids[:, :max_tokens].to(device)
    seq_len = input_ids.shape[1]
    if seq_len < 2:
        return dict(_EMPTY)

    budget = start_size + heavy_hitter_size + recent_size
    loss_fn = CrossEntropyLoss(reduction="none")
    past_key_values = None
    nlls, correct, token_times = [], [], []
    ttft = None
    steps = range(seq_len - 1)
    if show_progress:
        try:
            from tqdm import tqdm

            steps = tqdm(steps, desc="H2O-Attention Eval")
        except Exception:
            pass
    model.eval()
    vocab = model.config.vocab_size
    total_start = time.perf_counter()
    with torch.inference_mode():
        for idx in steps:
            token_start = time.perf_counter()
            outputs = model(input_ids[:, idx:idx + 1], past_key_values=past_key_values, use_cache=True,
                            output_attentions=True)
            logits = outputs.logits[:, -1, :].view(-1, vocab)
            target = input_ids[:, idx + 1:idx + 2].view(-1)
            nlls.append(float(loss_fn(logits, target).mean().item()))
            correct.append(float((torch.argmax(logits, dim=-1) == target).float().mean().item()))
            past_key_values = outputs.past_key_values
            attentions = outputs.attentions
            h2o_manager.update_attention_scores(attentions, skip_layers)
            if past_key_values is not None:
                kv_list = list(normalize_kv_cache(past_key_values))
                if kv_list and kv_list[0][0].size(2) > budget:
                    kept = h2o_attention_compress(kv_list, attention_scores=attentions, h2o_manager=h2o_manager,
                                                  start_size=start_size, heavy_hitter_size=heavy_hitter_size,
                                                  recent_size=recent_size, skip_layers=skip_layers)
                    past_key_values = to_dynamic_cache(kept)
                    h2o_manager.reset()
            token_time = time.perf_counter() - token_start
            token_times.append(token_time)
            if ttft is None:
                ttft = token_time
    total_time = time.perf_counter() - total_start

    num_tokens = len(nlls)
    lengths = [k.size(2) for k, _ in normalize_kv_cache(past_key_values)] if past_key_values is not None else []
    result = {
        "perplexity": math.exp(sum(nlls) / num_tokens),
        "accuracy": sum(correct) / num_tokens,
        "num_tokens": num_tokens,
        "final_cache_size": _final_cache_size(lengths, skip_layers),
        "ttft": ttft or 0.0,
        "tpot": sum(token_times[1:]) / (num_tokens - 1) if num_tokens > 1 else (ttft or 0.0),
        "throughput": num_tokens / total_time if total_time > 0 else 0.0,
        "total_time": total_time,
    }
    if return_nlls:
        result["nlls"] = nlls
        result["cache_lengths"] = lengths
    return result


def compare_h2o_methods(model, tokenizer=None, text: str = "", max_tokens: int = 2000,
                        heavy_hitter_sizes: List[int] = [32, 64, 128], skip_layers: List[int] = [0, 1],
                        device: Optional[torch.device] = None, **extra) -> List[Dict]:
    """Baseline, then for every heavy-hitter size (cache budget 512) H2O by key norm and H2O by attention score;
    each result carries ``method`` (reference evaluate_attention.py:231-330)."""
    from .methods import h2o_l2_compress

    if device is None:
        device = next(model.parameters()).device

    def report(label, res):
        res["method"] = label
        print(f"  {label}: PPL {res['perplexity']:.2f}, Acc {res['accuracy']:.2%}, cache {res['final_cache_size']}")
        return res

    results = [report("baseline", evaluate_with_compression(model, tokenizer, text, compress_fn=None,
                                                            max_tokens=max_tokens, device=device, **extra))]
    for hh in heavy_hitter_sizes:
        recent = 512 - 4 - hh
        results.append(report(f"h2o_l2_hh{hh}", evaluate_with_compression(
            model, tokenizer, text, compress_fn=h2o_l2_compress,
            compress_kwargs={"start_size": 4, "heavy_hitter_size": hh, "recent_size": recent},
            max_tokens=max_tokens, skip_layers=skip_layers, device=device, **extra)))
        results.append(report(f"h2o_attn_hh{hh}", evaluate_with_attention_compression(
            model, tokenizer, text, start_size=4, heavy_hitter_size=hh, recent_size=recent, max_tokens=max_tokens,
            skip_layers=skip_layers, device=device, **extra)))
    return results


__all__ = ["evaluate_with_attention_compression", "compare_h2o_methods"]
