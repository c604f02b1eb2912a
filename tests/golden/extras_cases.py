"""Inputs shared by make_extras_golden.py (CPU, real reference) and the tests: evict_for_space
(reference streaming_llm.py:114-170) and h2o_attention with a manager (h2o_attention.py:216-363)."""

import torch

EVICT_CASES = [
    # name, seq_lens, num_coming, start_size, recent_size, skip_layers
    ("evict_1", [600, 600, 600], 1, 4, 508, []),
    ("evict_chunk64", [600, 520, 449], 64, 4, 508, [0]),
    ("evict_coming_ge_recent", [700, 700], 600, 4, 508, []),     # effective_recent <= 0 -> falls back to recent
    ("evict_under_cap", [300, 511, 512], 1, 4, 508, []),
    ("evict_start0", [900], 16, 0, 128, []),
]

H2O_CASE = dict(layers=3, heads=4, dim=16, S=700, steps=3, start_size=4, heavy_hitter_size=32, recent_size=92,
                decay_factor=0.9, skip_layers=[0])


def evict_cache(seq_lens, seed=0):
    g = torch.Generator().manual_seed(seed)
    return [(torch.randn(1, 2, s, 16, generator=g), torch.arange(s, dtype=torch.float32).view(1, 1, s, 1).expand(1, 2, s, 16).contiguous())
            for s in seq_lens]


def h2o_inputs(step: int, seq_len: int):
    """Keys (random), values (= token position, recovers the kept rows) and softmax-like attention for one step."""
    c = H2O_CASE
    g = torch.Generator().manual_seed(100 + step)
    kv, attn = [], []
    for _ in range(c["layers"]):
        k = torch.randn(1, c["heads"], seq_len, c["dim"], generator=g)
        pos = torch.arange(seq_len, dtype=torch.float32).view(1, 1, seq_len, 1).expand(1, c["heads"], seq_len, c["dim"]).contiguous()
        kv.append((k, pos))
        attn.append(torch.softmax(2.0 * torch.randn(1, c["heads"], 1, seq_len, generator=g), dim=-1))
    return kv, attn
