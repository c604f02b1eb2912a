"""Golden-vector cases shared by ``make_golden.py`` (runs the real reference, build container only)
and the tests (which only read the committed ``.npz``).

Inputs are generated with numpy's PCG64 so that they can be regenerated bit-identically anywhere
without torch's RNG; bf16 rounding is done with the oracle's own round-to-nearest-even.
"""

from __future__ import annotations

import os
import sys

import numpy as np

_ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from oracle import kvc_oracle as O  # noqa: E402

GOLDEN_NPZ = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kvcompress_golden.npz")


def make_cache(seed: int, seq_lens, B: int, H: int, D: int, dtype: str, style: str = "spread"):
    """Synthetic cache: list of (K, V) numpy arrays in oracle storage format.

    spread: K = randn * exp(0.35 * randn_per_token), first 4 tokens * 0.1
    randn : plain randn (the reference's own fixture style, test_recent_only.py:34-35)
    ties  : every K row is one of 16 codebook rows -> many exactly equal norms
    """
    rng = np.random.default_rng(seed)
    layers = []
    for S in seq_lens:
        if style == "ties":
            book = rng.standard_normal((16, D), dtype=np.float32)
            which = rng.integers(0, 16, size=(B, H, S))
            k = book[which]
        else:
            k = rng.standard_normal((B, H, S, D), dtype=np.float32)
            if style == "spread":
                scale = np.exp(np.float32(0.35) * rng.standard_normal((B, H, S, 1), dtype=np.float32))
                scale[:, :, :4] *= np.float32(0.1)
                k = (k * scale).astype(np.float32)
        v = rng.standard_normal((B, H, S, D), dtype=np.float32)
        layers.append((O.store(k, dtype), O.store(v, dtype)))
    return layers


def _case(name, method, kwargs, seq_lens, dtype="f32", style="spread", B=2, H=3, D=16, seed=None):
    return dict(name=name, method=method, kwargs=kwargs, seq_lens=list(seq_lens), dtype=dtype, style=style,
                B=B, H=H, D=D, seed=seed)


def all_cases():
    """Appendix A presets of SURVEY.md (scripts/benchmark.py:421-511, README.md:141-155), the
    BASELINE.json configs at reduced shape, branch boundaries and the pinned quirks (§8a)."""
    cases = []
    S4 = [1300, 1300, 1300, 1300]

    def add(name, method, kwargs, seq_lens=S4, **kw):
        cases.append(_case(name, method, kwargs, seq_lens, **kw))

    # --- presets, fp32 + bf16, three data styles
    presets = [
        ("recent_512", "recent_only", dict(window_size=512, skip_layers=[0, 1])),
        ("streaming_512", "streaming_llm", dict(start_size=4, recent_size=508, skip_layers=[0, 1])),
        ("streaming_default_skip", "streaming_llm", dict(start_size=4, recent_size=508)),
        ("h2o_512", "h2o_l2", dict(start_size=4, heavy_hitter_size=64, recent_size=444, skip_layers=[0, 1])),
        ("h2o_1024", "h2o_l2", dict(start_size=4, heavy_hitter_size=128, recent_size=892)),
        ("snapkv_512", "snapkv_lite", dict(observation_window=32, keep_size=512, skip_layers=[0, 1])),
        ("snapkv_w16", "snapkv_lite", dict(observation_window=16, keep_size=512)),
        ("snapkv_w64_1024", "snapkv_lite", dict(observation_window=64, keep_size=1024)),
        ("snapkv_nopool", "snapkv_lite", dict(observation_window=32, keep_size=512, pooling_kernel=1)),
        ("snapkv_pool4", "snapkv_lite", dict(observation_window=32, keep_size=512, pooling_kernel=4)),
        ("snapkv_pool7", "snapkv_lite", dict(observation_window=32, keep_size=256, pooling_kernel=7)),
        ("pyramid_512", "pyramid_kv", dict(base_size=512, layer_decay=0.9, min_size=64, skip_layers=[0, 1])),
        ("pyramid_256_linear", "pyramid_kv", dict(base_size=256, min_size=64, profile="linear")),
        ("pyramid_256_constant", "pyramid_kv", dict(base_size=256, profile="constant")),
        ("adaptive_512", "adaptive_l2", dict(target_size=512, soft_limit=256, hard_limit=1024, skip_layers=[0, 1])),
        ("adaptive_256", "adaptive_l2", dict(target_size=256)),
        ("fix_512_kr05", "fix_size_l2", dict(fix_kv_size=512, strategy="keep_low", keep_ratio=0.5, skip_layers=[0, 1])),
        ("fix_512_kr02", "fix_size_l2", dict(fix_kv_size=512, strategy="keep_low", keep_ratio=0.2, skip_layers=[0, 1])),
        ("fix_256_high", "fix_size_l2", dict(fix_kv_size=256, strategy="keep_high", keep_ratio=0.3, skip_layers=[])),
        ("fix_default", "fix_size_l2", dict()),
        ("l2_08_100", "l2_compress", dict(keep_ratio=0.8, prune_after=100, skip_layers=[0, 1])),
        ("l2_05_1000", "l2_compress", dict(keep_ratio=0.5, prune_after=1000)),
        ("l2_03_100", "l2_compress", dict(keep_ratio=0.3, prune_after=100, skip_layers=[])),
    ]
    for dtype in ("f32", "bf16"):
        for name, method, kwargs in presets:
            add(f"{name}/{dtype}/spread", method, kwargs, dtype=dtype, style="spread")
    for name, method, kwargs in presets:
        if method not in ("recent_only", "streaming_llm"):
            add(f"{name}/bf16/ties", method, kwargs, dtype="bf16", style="ties")
            add(f"{name}/f32/randn", method, kwargs, dtype="f32", style="randn", seq_lens=[1300, 1300, 700])
    add("h2o_512/f16/spread", "h2o_l2", dict(start_size=4, heavy_hitter_size=64, recent_size=444), dtype="f16")
    add("snapkv_512/f16/spread", "snapkv_lite", dict(observation_window=32, keep_size=512), dtype="f16")
    add("fix_512_kr02/f16/ties", "fix_size_l2", dict(fix_kv_size=512, keep_ratio=0.2, skip_layers=[]), dtype="f16",
        style="ties")

    # --- sequence-length boundaries (identity below / at cap, steady state cap+1), ragged layers
    ragged = [100, 511, 512, 513, 640, 2000]
    for name, method, kwargs in presets:
        if name in ("streaming_default_skip", "h2o_512", "snapkv_512", "pyramid_512", "adaptive_512", "fix_512_kr02",
                    "recent_512", "l2_08_100"):
            kw = dict(kwargs)
            kw["skip_layers"] = []
            add(f"{name}/f32/ragged", method, kw, seq_lens=ragged, B=1, H=2, D=8)
    for S in (256, 257, 300, 640, 1023, 1024, 1025):
        add(f"adaptive_default/S{S}", "adaptive_l2", dict(), seq_lens=[S, S], B=1, H=2, D=8)
    add("adaptive_soft4/S6", "adaptive_l2", dict(target_size=4, soft_limit=2, hard_limit=64), seq_lens=[6, 9, 33],
        B=1, H=2, D=8)
    add("pyramid_tiny_budgets", "pyramid_kv", dict(base_size=12, layer_decay=0.5, min_size=1), seq_lens=[40] * 6,
        B=1, H=2, D=8)

    # --- quirks pinned by SURVEY.md §8a
    add("quirk/streaming_recent0", "streaming_llm", dict(start_size=4, recent_size=0), seq_lens=[600, 3], B=1, H=2, D=8)
    add("quirk/fix_keep_ratio_1", "fix_size_l2", dict(fix_kv_size=512, keep_ratio=1.0, skip_layers=[]),
        seq_lens=[700, 400], B=1, H=2, D=8)
    add("quirk/fix_zone_small", "fix_size_l2", dict(fix_kv_size=8, keep_ratio=0.9, skip_layers=[]), seq_lens=[9, 30],
        B=1, H=2, D=8)
    add("quirk/h2o_hh0", "h2o_l2", dict(start_size=4, heavy_hitter_size=0, recent_size=28), seq_lens=[100], B=1, H=2,
        D=8)
    add("quirk/snapkv_keep_le_window", "snapkv_lite", dict(observation_window=32, keep_size=16), seq_lens=[100],
        B=1, H=2, D=8)
    add("quirk/snapkv_short_prefix", "snapkv_lite", dict(observation_window=32, keep_size=33, pooling_kernel=5),
        seq_lens=[35, 36, 40], B=1, H=2, D=8)
    add("quirk/l2_keep_all", "l2_compress", dict(keep_ratio=1.0, prune_after=10), seq_lens=[50, 50, 50], B=1, H=2, D=8)
    add("quirk/l2_tiny_ratio", "l2_compress", dict(keep_ratio=0.001, prune_after=10, skip_layers=[]), seq_lens=[50, 2000],
        B=1, H=2, D=8)
    add("quirk/empty_cache", "streaming_llm", dict(), seq_lens=[], B=1, H=2, D=8)

    # --- the reference's own fixture shape (test_recent_only.py:25-36)
    for w in (256, 512, 1024):
        add(f"fixture/recent_only_w{w}", "recent_only", dict(window_size=w, skip_layers=[0, 1]), seq_lens=[1000] * 4,
            B=1, H=8, D=64, style="randn")

    # --- BASELINE.json configs at reduced batch/heads (full head_dim, full-ish context)
    add("c1/l2_08_pythia", "l2_compress", dict(keep_ratio=0.8, prune_after=1000, skip_layers=[0, 1]),
        seq_lens=[2048] * 3, B=1, H=2, D=80, dtype="f32")
    add("c2/streaming_pythia", "streaming_llm", dict(start_size=4, recent_size=508), seq_lens=[4096, 513], B=1, H=2,
        D=80, dtype="bf16")
    add("c2/fix_pythia", "fix_size_l2", dict(fix_kv_size=512, keep_ratio=0.2, strategy="keep_low", skip_layers=[]),
        seq_lens=[4096, 513], B=1, H=2, D=80, dtype="bf16")
    add("c3/h2o_pythia", "h2o_l2", dict(start_size=4, heavy_hitter_size=64, recent_size=444), seq_lens=[8192], B=1,
        H=2, D=80, dtype="bf16")
    add("c4/snapkv_llama", "snapkv_lite", dict(observation_window=32, keep_size=512), seq_lens=[32768], B=1, H=1,
        D=128, dtype="bf16")
    add("c5/pyramid_llama", "pyramid_kv", dict(base_size=512), seq_lens=[8192, 8192, 8192], B=1, H=1, D=128,
        dtype="bf16")
    add("c5/adaptive_llama", "adaptive_l2", dict(target_size=512), seq_lens=[8192], B=1, H=1, D=128, dtype="bf16")

    for i, c in enumerate(cases):
        if c["seed"] is None:
            c["seed"] = 1234 + 7 * i
    names = [c["name"] for c in cases]
    assert len(set(names)) == len(names), "duplicate case names"
    return cases


def case_cache(case):
    return make_cache(case["seed"], case["seq_lens"], case["B"], case["H"], case["D"], case["dtype"], case["style"])
