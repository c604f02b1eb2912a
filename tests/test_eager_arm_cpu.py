"""The bench-only eager-torch comparison arm (scripts/torch_eager_methods.py) must be a faithful restatement of the
reference's op sequence: on fp32 data (no ties) it keeps exactly the rows the golden-pinned oracle keeps."""

import os
import sys

import numpy as np
import pytest
import torch

import cases
from oracle import kvc_oracle as O

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "scripts"))
from torch_eager_methods import eager_fn  # noqa: E402

PRESETS = [
    ("streaming_llm", dict(start_size=4, recent_size=508)),
    ("h2o_l2", dict(start_size=4, heavy_hitter_size=64, recent_size=444)),
    ("snapkv_lite", dict(observation_window=32, keep_size=512)),
    ("pyramid_kv", dict(base_size=512, layer_decay=0.9, min_size=64)),
    ("adaptive_l2", dict(target_size=512, soft_limit=256, hard_limit=1024)),
    ("fix_size_l2", dict(fix_kv_size=512, strategy="keep_low", keep_ratio=0.5)),
    ("fix_size_l2", dict(fix_kv_size=256, strategy="keep_high", keep_ratio=0.3)),
    ("l2_compress", dict(keep_ratio=0.8, prune_after=100)),
    ("recent_only", dict(window_size=512)),
]


@pytest.mark.parametrize("method,kwargs", PRESETS, ids=[f"{m}-{i}" for i, (m, _) in enumerate(PRESETS)])
def test_eager_arm_keeps_the_oracle_rows(method, kwargs):
    case = cases._case("eager", method, kwargs, [1300, 1300, 700], dtype="f32", style="randn", B=1, H=2, D=16, seed=3)
    layers = cases.case_cache(case)
    kv = []
    for K, _ in layers:
        k = torch.from_numpy(K.copy())
        pos = torch.arange(k.size(2), dtype=torch.float32).view(1, 1, -1, 1).expand(k.shape).contiguous()
        kv.append((k, pos))
    out = eager_fn(method)(kv, skip_layers=[0], **kwargs)
    results = O.METHODS[method](layers, "f32", skip_layers=[0], **kwargs)
    for (k_in, _), (k_out, v_out), res in zip(kv, out, results):
        if res.untouched:
            assert k_out is k_in
            continue
        assert np.array_equal(v_out[..., 0].numpy().astype(np.int64), res.rows)
