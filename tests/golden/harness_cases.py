"""Shared by make_harness_golden.py (CPU, real reference) and tests/test_gpu_harness.py (B200)."""

import torch

MODEL_KW = dict(vocab_size=512, hidden_size=256, num_hidden_layers=4, num_attention_heads=4, intermediate_size=512,
                max_position_embeddings=2048, rotary_pct=0.25)
TOKENS = 110
SKIP = [0, 1]
CASES = [
    ("streaming", "streaming_llm", dict(start_size=4, recent_size=28)),
    ("h2o_l2", "h2o_l2", dict(start_size=4, heavy_hitter_size=8, recent_size=20)),
    ("fix_size", "fix_size_l2", dict(fix_kv_size=32, keep_ratio=0.25)),
    ("fix_size_high", "fix_size_l2", dict(fix_kv_size=32, keep_ratio=0.5, strategy="keep_high")),
    ("snapkv", "snapkv_lite", dict(observation_window=8, keep_size=32, pooling_kernel=5)),
    ("pyramid", "pyramid_kv", dict(base_size=32, layer_decay=0.8, min_size=8)),
    ("adaptive", "adaptive_l2", dict(target_size=32, soft_limit=16, hard_limit=48)),
    ("l2", "l2_compress", dict(keep_ratio=0.9, prune_after=30)),
    ("recent_only", "recent_only", dict(window_size=24)),
]


ATTN_KW = dict(start_size=4, heavy_hitter_size=8, recent_size=20)   # the attention-score harness (evaluate_attention.py)


def tiny_model_and_ids(device="cpu", attn_implementation=None):
    from transformers import GPTNeoXConfig, GPTNeoXForCausalLM

    torch.manual_seed(0)
    extra = {"attn_implementation": attn_implementation} if attn_implementation else {}
    model = GPTNeoXForCausalLM(GPTNeoXConfig(**MODEL_KW, **extra)).eval()  # fp32, head_dim 64 -> 256-byte rows
    ids = torch.randint(0, MODEL_KW["vocab_size"], (1, TOKENS), generator=torch.Generator().manual_seed(1))
    return model.to(device), ids
