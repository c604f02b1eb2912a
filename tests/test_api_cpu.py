"""Drop-in surface on CPU: registry, signatures, no-op / view semantics, loud failure without CUDA."""

import inspect

import pytest
import torch

import kvcompress
from kvcompress import methods


def make_kv(L=4, B=1, H=8, S=1000, D=64, dtype=torch.float32):
    return [(torch.randn(B, H, S, D).to(dtype), torch.randn(B, H, S, D).to(dtype)) for _ in range(L)]


def test_registry_names_and_order():
    # reference methods/__init__.py:21-33
    assert kvcompress.list_methods() == ["l2_compress", "fix_size_l2", "streaming_llm", "recent_only", "h2o_l2",
                                         "h2o_attention", "snapkv_lite", "pyramid_kv", "adaptive_l2"]
    for name in kvcompress.list_methods():
        assert callable(kvcompress.get_compress_fn(name))
    with pytest.raises(ValueError, match=r"Unknown method: nope\. Available: \['l2_compress'"):
        kvcompress.get_compress_fn("nope")
    kvcompress.register_method("mine", lambda kv, **kw: kv)
    assert "mine" in kvcompress.list_methods()
    del kvcompress.COMPRESS_METHODS["mine"]
    assert kvcompress.__version__ == "2.0.0"


def test_signatures_and_defaults_match_reference():
    want = {
        "l2_compress": dict(keep_ratio=1.0, prune_after=1000, skip_layers=[0, 1]),
        "fix_size_l2": dict(fix_kv_size=1024, keep_ratio=0.0, strategy="keep_low", skip_layers=[0, 1]),
        "streaming_llm": dict(start_size=4, recent_size=508, skip_layers=[]),
        "recent_only": dict(window_size=512, skip_layers=[0, 1]),
        "h2o_l2": dict(start_size=4, heavy_hitter_size=64, recent_size=444, skip_layers=[]),
        "snapkv_lite": dict(observation_window=32, keep_size=512, pooling_kernel=5, skip_layers=[]),
        "pyramid_kv": dict(base_size=512, layer_decay=0.9, min_size=64, profile="exponential", skip_layers=[]),
        "adaptive_l2": dict(target_size=512, soft_limit=256, hard_limit=1024, keep_ratio_min=0.3, keep_ratio_max=0.9,
                            skip_layers=[]),
    }
    for name, defaults in want.items():
        sig = inspect.signature(kvcompress.get_compress_fn(name))
        params = list(sig.parameters.values())
        assert params[0].name == "past_key_values"
        assert params[-1].kind is inspect.Parameter.VAR_KEYWORD
        got = {p.name: p.default for p in params[1:-1]}
        assert got == defaults, name
        assert list(got) == list(defaults), name  # positional order too
    # every hot-path function is importable from the top level and from .methods
    for fn in ("l2_compress", "fix_size_l2_compress", "streaming_llm_compress", "h2o_l2_compress",
               "snapkv_lite_compress", "pyramid_kv_compress", "adaptive_l2_compress", "recent_only_compress"):
        assert getattr(kvcompress, fn) is getattr(methods, fn)
    from kvcompress.methods.base import CompressFn
    assert isinstance(kvcompress.l2_compress, CompressFn)


def test_recent_only_fixture_of_reference():
    """The reference's only test (test_recent_only.py:25-56): shapes after recent_only_compress."""
    kv = make_kv()
    for window, want in ((256, [1000, 1000, 256, 256]), (512, [1000, 1000, 512, 512]), (1024, [1000] * 4)):
        out = kvcompress.recent_only_compress(kv, window_size=window, skip_layers=[0, 1])
        assert [k.size(2) for k, _ in out] == want
        for (ki, vi), (ko, vo) in zip(kv, out):
            if ko.size(2) == 1000:
                assert ko is ki and vo is vi  # untouched layers are the same tensor objects
            else:
                assert ko._is_view() and not ko.is_contiguous()
                assert torch.equal(ko, ki[:, :, -window:, :]) and torch.equal(vo, vi[:, :, -window:, :])
                assert ko.data_ptr() == ki[:, :, -window:, :].data_ptr()  # aliases the input


def test_noop_calls_return_same_objects_and_do_not_mutate_input():
    kv = make_kv(L=3, S=300, H=2, D=16)
    snapshot = list(kv)
    for name, kwargs in (("l2_compress", dict(keep_ratio=0.5)), ("fix_size_l2", {}), ("streaming_llm", {}),
                         ("h2o_l2", {}), ("snapkv_lite", {}), ("pyramid_kv", dict(base_size=2000)),
                         ("adaptive_l2", dict(soft_limit=300)), ("recent_only", {}), ("h2o_attention", {})):
        out = kvcompress.get_compress_fn(name)(kv, **kwargs)
        assert out is not kv and isinstance(out, list)
        assert all(a[0] is b[0] and a[1] is b[1] for a, b in zip(out, kv)), name
        assert kv == snapshot
    assert kvcompress.streaming_llm_compress([]) == []


def test_ignores_unknown_kwargs_like_reference():
    kv = make_kv(L=2, S=50, H=2, D=16)
    kvcompress.l2_compress(kv, keep_ratio=0.5, some_future_flag=True)


def test_view_paths_run_on_cpu():
    kv = make_kv(L=2, S=700, H=2, D=16)
    out = kvcompress.fix_size_l2_compress(kv, fix_kv_size=512, keep_ratio=1.0, skip_layers=[])
    assert [k.size(2) for k, _ in out] == [512, 512] and out[0][0]._is_view()
    out = kvcompress.snapkv_lite_compress(kv, observation_window=32, keep_size=16)
    assert [k.size(2) for k, _ in out] == [32, 32] and torch.equal(out[1][1], kv[1][1][:, :, -32:])


def test_no_cpu_fallback():
    kv = make_kv(L=1, S=700, H=2, D=16)
    for fn, kwargs in ((kvcompress.streaming_llm_compress, {}), (kvcompress.h2o_l2_compress, {}),
                       (kvcompress.l2_compress, dict(keep_ratio=0.5, prune_after=10, skip_layers=[]))):
        with pytest.raises(RuntimeError, match="no CPU path"):
            fn(kv, **kwargs)


def test_normalize_tolerates_transformers5_dynamic_cache():
    kv = make_kv(L=2, S=10, H=2, D=16)
    cache = kvcompress.to_dynamic_cache(kv)
    back = kvcompress.normalize_kv_cache(cache)
    assert len(back) == 2 and all(len(item) == 2 for item in back)
    assert torch.equal(back[1][0], kv[1][0])
    assert kvcompress.get_seq_len(cache) == 10 and kvcompress.get_seq_len(kv, 5) == 0
    info = kvcompress.get_cache_info(kv)
    assert info["num_layers"] == 2 and info["seq_lengths"] == [10, 10]
    assert abs(kvcompress.get_cache_size_mb(kv) - 2 * 2 * 2 * 10 * 16 * 4 / 2 ** 20) < 1e-12
    # the reference's own call convention: compress_fn(list(normalize_kv_cache(cache)), skip_layers=..., **kw)
    out = kvcompress.recent_only_compress(cache, window_size=4, skip_layers=[0])
    assert [k.size(2) for k, _ in out] == [10, 4]


def test_h2o_manager_bookkeeping():
    m = kvcompress.H2OAttentionManager(start_size=1, heavy_hitter_size=2, recent_size=2, decay_factor=0.5)
    attn = torch.zeros(1, 2, 1, 8)
    attn[0, :, 0, 3] = 1.0
    m.update_attention_scores((attn,))
    attn2 = torch.zeros(1, 2, 1, 9)
    attn2[0, :, 0, 5] = 0.75
    m.update_attention_scores((attn2,))
    acc = m.accumulated_attention[0]
    assert acc.shape == (1, 2, 9) and acc[0, 0, 3] == 0.5 and acc[0, 0, 5] == 0.75
    idx = m.get_heavy_hitter_indices(0, 9)   # middle = rows [1, 7): relative positions of rows 3 and 5
    assert idx.tolist() == [2, 4]
    assert m.get_heavy_hitter_indices(7, 30).tolist() == list(range(0, 27, 13))[:2]


def test_slab_cache_planner_table_matches_function_signatures():
    """Every in-place method resolves its planner arguments from the drop-in function's own defaults."""
    import inspect

    from kvcompress import _planner as P
    from kvcompress import slab_cache

    for name, (planner, names) in slab_cache._PLANNERS.items():
        defaults = slab_cache._method_defaults(name)
        assert set(names) <= set(defaults), (name, names, defaults)
        assert "skip_layers" in defaults
        params = list(inspect.signature(planner).parameters)
        want = params[1:params.index("skip_layers")]
        assert len(want) == len(names), (name, want, names)
    plans = [P.LayerPlan(P.KEEP, 10), P.LayerPlan(P.VIEW, 10, view_n=4), P.LayerPlan(P.VIEW, 10, view_n=0)]
    moved = slab_cache._in_place(plans)
    assert moved[0].kind == P.KEEP
    assert (moved[1].kind, moved[1].tail, moved[1].out_len) == (P.GATHER, 4, 4)
    assert (moved[2].kind, moved[2].tail) == (P.GATHER, 10)  # x[:, :, -0:] is the whole tensor
    with pytest.raises(RuntimeError, match="no CPU path"):
        slab_cache.KVSlabCache(1, 1, 1, 80, 16, device="cpu")


def test_integration_route2_snippet_registers_into_a_reference_style_registry(monkeypatch):
    """INTEGRATION.md route 2, executed as written: the fenced block a maintainer pastes into the reference's
    methods/__init__.py must import the package under a private name (relative imports included) and register the
    eight functions; with a bad KVCOMPRESS_B200 it must fall through silently."""
    import os
    import re
    import sys

    root = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    block = re.search(r"```python\n(try:\s+# B200 fast path.*?)```", text, flags=re.S).group(1)
    registry = {}
    scope = {"register_method": lambda name, fn: registry.__setitem__(name, fn)}
    monkeypatch.setenv("KVCOMPRESS_B200", os.path.join(root, "cs3602-llm-inference-acceleration_b200"))
    monkeypatch.delitem(sys.modules, "kvcompress_b200", raising=False)
    exec(block, scope)
    assert sorted(registry) == sorted(["l2_compress", "fix_size_l2", "streaming_llm", "h2o_l2", "snapkv_lite",
                                       "pyramid_kv", "adaptive_l2", "recent_only"])
    assert registry["h2o_l2"].__module__.startswith("kvcompress_b200.")
    sys.modules.pop("kvcompress_b200", None)
    for name in [m for m in sys.modules if m.startswith("kvcompress_b200.")]:
        sys.modules.pop(name, None)
    registry.clear()
    monkeypatch.setenv("KVCOMPRESS_B200", "/nonexistent")
    exec(block, scope)            # FileNotFoundError is an OSError: swallowed, registry untouched
    assert registry == {}
    monkeypatch.delenv("KVCOMPRESS_B200")
    exec(block, scope)            # KeyError: swallowed
    assert registry == {}


def test_route1_exports_the_reference_top_level_surface():
    """Everything `from kvcompress import ...` offers in the reference (kvcompress/__init__.py:33-97)."""
    import kvcompress

    for name in ("l2_compress", "fix_size_l2_compress", "streaming_llm_compress", "get_compress_fn", "list_methods",
                 "register_method", "COMPRESS_METHODS", "evaluate_with_compression", "evaluate_baseline",
                 "compare_methods", "benchmark", "measure_generation_metrics", "run_benchmark_suite",
                 "print_benchmark_summary", "to_dynamic_cache", "normalize_kv_cache", "get_cache_size_mb",
                 "get_cache_info", "get_seq_len"):
        assert hasattr(kvcompress, name) and name in kvcompress.__all__, name
