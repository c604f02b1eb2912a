"""Host planner (product) vs the golden vectors of the real reference and vs the oracle."""

import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

import cases
from kvcompress import _planner as P
from oracle import kvc_oracle as O

ALL = cases.all_cases()


def plan_for(method, seq_lens, kwargs):
    kw = dict(kwargs)
    if method == "l2_compress":
        return P.plan_l2(seq_lens, kw.get("keep_ratio", 1.0), kw.get("prune_after", 1000), kw.get("skip_layers", [0, 1]))
    if method == "fix_size_l2":
        return P.plan_fix_size(seq_lens, kw.get("fix_kv_size", 1024), kw.get("keep_ratio", 0.0),
                               kw.get("strategy", "keep_low"), kw.get("skip_layers", [0, 1]))
    if method == "streaming_llm":
        return P.plan_streaming(seq_lens, kw.get("start_size", 4), kw.get("recent_size", 508), kw.get("skip_layers", []))
    if method == "recent_only":
        return P.plan_recent_only(seq_lens, kw.get("window_size", 512), kw.get("skip_layers", [0, 1]))
    if method == "h2o_l2":
        return P.plan_h2o(seq_lens, kw.get("start_size", 4), kw.get("heavy_hitter_size", 64), kw.get("recent_size", 444),
                          kw.get("skip_layers", []))
    if method == "snapkv_lite":
        return P.plan_snapkv(seq_lens, kw.get("observation_window", 32), kw.get("keep_size", 512),
                             kw.get("pooling_kernel", 5), kw.get("skip_layers", []))
    if method == "pyramid_kv":
        return P.plan_pyramid(seq_lens, kw.get("base_size", 512), kw.get("layer_decay", 0.9), kw.get("min_size", 64),
                              kw.get("profile", "exponential"), kw.get("skip_layers", []))
    if method == "adaptive_l2":
        return P.plan_adaptive(seq_lens, kw.get("target_size", 512), kw.get("soft_limit", 256), kw.get("hard_limit", 1024),
                               kw.get("keep_ratio_min", 0.3), kw.get("keep_ratio_max", 0.9), kw.get("skip_layers", []))
    raise KeyError(method)


def plan_rows(plan):
    """Kept rows implied by a plan when nothing is selected (sink + tail only)."""
    S = plan.seq_len
    if plan.kind == P.VIEW:
        return np.arange(S)[-plan.view_n:]
    return np.concatenate([np.arange(plan.sink), np.arange(S - plan.tail, S)])


@pytest.mark.parametrize("case", ALL, ids=[c["name"] for c in ALL])
def test_plans_match_reference(case, golden):
    data, manifest = golden
    meta = manifest[case["name"]]
    plans = plan_for(case["method"], case["seq_lens"], case["kwargs"])
    assert [p.out_len for p in plans] == meta["lengths"]
    assert [p.kind == P.KEEP for p in plans] == meta["untouched"]
    assert [p.kind == P.VIEW for p in plans] == meta["view"]
    for li, p in enumerate(plans):
        if p.kind == P.KEEP:
            continue
        ref_rows = data[f"{case['name']}|L{li}"].astype(np.int64)
        if p.kind == P.VIEW or p.k_sel == 0:
            assert np.array_equal(ref_rows, np.broadcast_to(plan_rows(p), ref_rows.shape))
        else:
            # sinks, tail and the selection window agree with what the reference kept
            assert np.array_equal(ref_rows[..., :p.sink], np.broadcast_to(np.arange(p.sink), ref_rows[..., :p.sink].shape))
            tail = ref_rows[..., p.sink + p.k_sel:]
            assert np.array_equal(tail, np.broadcast_to(np.arange(p.seq_len - p.tail, p.seq_len), tail.shape))
            sel = ref_rows[..., p.sink:p.sink + p.k_sel]
            assert sel.min() >= p.sel_lo and sel.max() < p.sel_hi


def _oracle_lengths(method, seq_lens, kwargs):
    layers = [(np.zeros((1, 1, s, 4), np.float32) + np.arange(s, dtype=np.float32)[None, None, :, None],
               np.zeros((1, 1, s, 4), np.float32)) for s in seq_lens]
    res = O.METHODS[method](layers, "f32", **kwargs)
    return O.out_lengths(layers, res), [r.untouched for r in res], [(not r.untouched) and r.is_view for r in res], res


def _assert_same(method, seq_lens, kwargs):
    plans = plan_for(method, seq_lens, kwargs)
    lengths, untouched, view, res = _oracle_lengths(method, seq_lens, kwargs)
    assert [p.out_len for p in plans] == lengths
    assert [p.kind == P.KEEP for p in plans] == untouched
    assert [p.kind == P.VIEW for p in plans] == view
    for p, r in zip(plans, res):
        if p.kind == P.GATHER:
            assert (p.sel_lo, p.sel_hi) == (r.region if r.k_sel else (p.sel_lo, p.sel_hi))
            assert p.k_sel == r.k_sel
            assert 0 <= p.sink <= p.seq_len and 0 <= p.tail <= p.seq_len
            if p.k_sel:
                assert 0 <= p.sel_lo < p.sel_hi <= p.seq_len and p.k_sel <= p.sel_hi - p.sel_lo


lens = st.lists(st.integers(1, 3000), min_size=1, max_size=6)
skips = st.lists(st.integers(0, 5), max_size=3)


@settings(max_examples=150, deadline=None)
@given(lens, st.floats(0.01, 1.2), st.integers(0, 2500), skips)
def test_prop_l2(seq_lens, keep_ratio, prune_after, skip):
    _assert_same("l2_compress", seq_lens, dict(keep_ratio=keep_ratio, prune_after=prune_after, skip_layers=skip))


@settings(max_examples=150, deadline=None)
@given(lens, st.integers(0, 2048), st.floats(0.0, 1.3), st.sampled_from(["keep_low", "keep_high"]), skips)
def test_prop_fix_size(seq_lens, size, keep_ratio, strategy, skip):
    _assert_same("fix_size_l2", seq_lens, dict(fix_kv_size=size, keep_ratio=keep_ratio, strategy=strategy, skip_layers=skip))


@settings(max_examples=150, deadline=None)
@given(lens, st.integers(0, 64), st.integers(0, 2048), skips)
def test_prop_streaming(seq_lens, start, recent, skip):
    _assert_same("streaming_llm", seq_lens, dict(start_size=start, recent_size=recent, skip_layers=skip))


@settings(max_examples=150, deadline=None)
@given(lens, st.integers(0, 16), st.integers(0, 300), st.integers(0, 1500), skips)
def test_prop_h2o(seq_lens, start, hh, recent, skip):
    _assert_same("h2o_l2", seq_lens, dict(start_size=start, heavy_hitter_size=hh, recent_size=recent, skip_layers=skip))


@settings(max_examples=150, deadline=None)
@given(lens, st.integers(0, 128), st.integers(0, 1500), st.integers(1, 9), skips)
def test_prop_snapkv(seq_lens, window, keep, kernel, skip):
    _assert_same("snapkv_lite", seq_lens, dict(observation_window=window, keep_size=keep, pooling_kernel=kernel, skip_layers=skip))


@settings(max_examples=150, deadline=None)
@given(lens, st.integers(0, 1500), st.floats(0.3, 1.0), st.integers(0, 200),
       st.sampled_from(["exponential", "linear", "constant"]), skips)
def test_prop_pyramid(seq_lens, base, decay, min_size, profile, skip):
    _assert_same("pyramid_kv", seq_lens, dict(base_size=base, layer_decay=decay, min_size=min_size, profile=profile, skip_layers=skip))


@settings(max_examples=200, deadline=None)
@given(lens, st.integers(0, 1500), st.integers(0, 600), st.integers(601, 2500), st.floats(0.05, 0.6),
       st.floats(0.6, 1.0), skips)
def test_prop_adaptive(seq_lens, target, soft, hard, rmin, rmax, skip):
    _assert_same("adaptive_l2", seq_lens, dict(target_size=target, soft_limit=soft, hard_limit=hard, keep_ratio_min=rmin,
                                                keep_ratio_max=rmax, skip_layers=skip))


@settings(max_examples=100, deadline=None)
@given(lens, st.integers(0, 2048), skips)
def test_prop_recent_only(seq_lens, window, skip):
    _assert_same("recent_only", seq_lens, dict(window_size=window, skip_layers=skip))


def test_pyramid_budgets_match_survey():
    sizes = P.pyramid_layer_sizes(32, 512, 0.9, 64, "exponential")
    assert sizes[:4] == [512, 460, 414, 373] and sizes[-12:] == [64] * 12 and sum(sizes) == 5256


def test_adaptive_steady_state_lengths():
    # SURVEY.md appendix B.8
    for S, want in ((257, 256), (1024, 307), (1025, 512)):
        assert P.plan_adaptive([S], 512, 256, 1024, 0.3, 0.9, [])[0].out_len == want


def test_algorithmic_bytes_match_baseline_md():
    """BASELINE.md §3 per-step byte counts, from the planner."""
    e = 2
    c2a = P.plan_streaming([4096] * 32, 4, 508, [])
    assert P.algorithmic_bytes(c2a, 32, 32, 80, e) == 10_737_418_240
    c2b = P.plan_fix_size([4096] * 32, 512, 0.2, "keep_low", [0, 1])
    assert round(P.algorithmic_bytes(c2b, 32, 32, 80, e) / 1e9, 3) == 29.698
    c2b_steady = P.plan_fix_size([513] * 32, 512, 0.2, "keep_low", [0, 1])
    assert round(P.algorithmic_bytes(c2b_steady, 32, 32, 80, e) / 1e9, 3) == 12.086
    c3 = P.plan_h2o([8192] * 32, 4, 64, 444, [])
    assert round(P.algorithmic_bytes(c3, 256, 32, 80, e) / 1e9, 1) == 410.7
    c4 = P.plan_snapkv([32768] * 32, 32, 512, 5, [])
    assert round(P.algorithmic_bytes(c4, 16, 8, 128, e) / 1e9, 3) == 36.474
    c5a = P.plan_pyramid([32768] * 32, 512, 0.9, 64, "exponential", [])
    assert round(P.algorithmic_bytes(c5a, 64, 8, 128, e) / 1e9, 1) == 139.8
    c5b = P.plan_adaptive([32768] * 32, 512, 256, 1024, 0.3, 0.9, [])
    assert round(P.algorithmic_bytes(c5b, 64, 8, 128, e) / 1e9, 1) == 144.9
    c1 = P.plan_l2([2048] * 32, 0.8, 1000, [0, 1])
    assert round(P.algorithmic_bytes(c1, 1, 32, 80, 4) / 1e9, 3) == 2.643


def test_unknown_strategy_raises_like_reference():
    with pytest.raises(ValueError, match="Unknown strategy: nope"):
        P.plan_fix_size([2000], 512, 0.2, "nope", [])
    # ...but only when a layer actually reaches the strategy switch (fix_size_l2.py:99-126)
    assert P.plan_fix_size([100], 512, 0.2, "nope", [])[0].kind == P.KEEP
