"""-m gpu: the caller loop (kvcompress/evaluate.py) on a tiny random-weight GPT-NeoX —
our compress functions (DynamicCache path, the reference's loop shape) and the in-place KVSlabCache path
against golden NLLs produced with the REAL reference's functions on CPU (tests/golden/make_harness_golden.py)."""

import json
import os

import pytest
import torch

import harness_cases as H
import kvcompress
from kvcompress import _engine
from kvcompress.benchmark import measure_generation_metrics
from kvcompress.evaluate import evaluate_with_compression

pytestmark = pytest.mark.gpu

GOLDEN = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "harness_golden.json")))


@pytest.fixture(scope="module")
def model_and_ids():
    model, ids = H.tiny_model_and_ids("cuda")
    return model, ids.cuda()


def test_baseline_matches_golden(model_and_ids):
    model, ids = model_and_ids
    r = evaluate_with_compression(model, input_ids=ids, compress_fn=None, show_progress=False, return_nlls=True)
    want = GOLDEN["cases"]["baseline"]
    assert r["cache_lengths"] == want["lengths"]
    assert max(abs(a - b) for a, b in zip(r["nlls"], want["nlls"])) < 2e-3  # fp32 GPU vs CPU matmul order


@pytest.mark.parametrize("name,method,kwargs", H.CASES, ids=[c[0] for c in H.CASES])
def test_loop_matches_reference_golden(model_and_ids, name, method, kwargs):
    model, ids = model_and_ids
    fn = kvcompress.get_compress_fn(method)
    want = GOLDEN["cases"][name]
    n0 = _engine.launch_count()
    dyn = evaluate_with_compression(model, input_ids=ids, compress_fn=fn, compress_kwargs=kwargs, skip_layers=H.SKIP,
                                    show_progress=False, return_nlls=True, cache="dynamic")
    assert _engine.launch_count() > n0 or method == "recent_only", "the sm_100a library must have run"
    slab = evaluate_with_compression(model, input_ids=ids, compress_fn=fn, compress_kwargs=kwargs, skip_layers=H.SKIP,
                                     show_progress=False, return_nlls=True, cache="slab")
    for got in (dyn, slab):
        assert got["cache_lengths"] == want["lengths"], name
        assert got["final_cache_size"] == want["final_cache_size"]
        assert got["num_tokens"] == H.TOKENS - 1
        # same kept rows every step -> same logits up to fp32 summation order (GPU vs the CPU golden)
        assert max(abs(a - b) for a, b in zip(got["nlls"], want["nlls"])) < 5e-3, name
        assert abs(got["perplexity"] / want["perplexity"] - 1) < 1e-3
    assert max(abs(a - b) for a, b in zip(dyn["nlls"], slab["nlls"])) < 1e-3  # the two cache paths agree


def test_generation_metrics_both_cache_paths(model_and_ids):
    model, ids = model_and_ids
    fn = kvcompress.get_compress_fn("h2o_l2")
    kw = dict(start_size=4, heavy_hitter_size=8, recent_size=20)
    for cache in ("dynamic", "slab"):
        r = measure_generation_metrics(model, input_ids=ids[:, :64], compress_fn=fn, compress_kwargs=kw,
                                       max_new_tokens=24, skip_layers=H.SKIP, cache=cache)
        assert r["num_tokens"] == 24 and r["input_length"] == 64 and r["ttft"] > 0 and r["tpot"] > 0


def test_attention_score_harness_matches_reference_golden():
    """kvcompress/evaluate_attention.py with OUR h2o_attention_compress + manager on the B200 against the loop driven
    with the REAL reference's on CPU (tests/golden/make_harness_golden.py): same kept rows every step -> same NLLs."""
    from kvcompress.evaluate_attention import compare_h2o_methods, evaluate_with_attention_compression

    model, ids = H.tiny_model_and_ids("cuda", attn_implementation="eager")
    ids = ids.cuda()
    want = GOLDEN["cases"]["h2o_attention"]
    n0 = _engine.launch_count()
    got = evaluate_with_attention_compression(model, input_ids=ids, skip_layers=[], show_progress=False,
                                              return_nlls=True, **H.ATTN_KW)
    assert _engine.launch_count() > n0, "the sm_100a library must have run"
    assert got["cache_lengths"] == want["lengths"] and got["final_cache_size"] == want["final_cache_size"]
    assert got["num_tokens"] == H.TOKENS - 1
    assert max(abs(a - b) for a, b in zip(got["nlls"], want["nlls"])) < 5e-3
    assert abs(got["perplexity"] / want["perplexity"] - 1) < 1e-3
    rows = compare_h2o_methods(model, input_ids=ids[:, :48], max_tokens=48, heavy_hitter_sizes=[8], skip_layers=[],
                               show_progress=False)
    assert [r["method"] for r in rows] == ["baseline", "h2o_l2_hh8", "h2o_attn_hh8"]
