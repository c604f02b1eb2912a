"""Generate tests/golden/harness_golden.json: the decode loop of kvcompress/evaluate.py driven with the
REAL reference's compress functions on CPU (build container only; /root/reference is absent on the GPU box).

    python tests/golden/make_harness_golden.py

A tiny random-weight GPT-NeoX (weights from torch.manual_seed, reproducible wherever this torch runs) is
evaluated token by token with each method applied after every step, exactly the reference's loop shape
(evaluate.py:121-166).  Stored per method: per-token NLLs, final per-layer cache lengths, PPL.
"""

import importlib.util
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.normpath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "cs3602-llm-inference-acceleration_b200"))
sys.path.insert(0, HERE)
sys.dont_write_bytecode = True

import harness_cases as H  # noqa: E402
from kvcompress.evaluate import evaluate_with_compression  # noqa: E402  (the loop under test)

spec = importlib.util.spec_from_file_location(
    "kvcompress_ref", "/root/reference/kvcompress/__init__.py",
    submodule_search_locations=["/root/reference/kvcompress"])
ref = importlib.util.module_from_spec(spec)
sys.modules["kvcompress_ref"] = ref
spec.loader.exec_module(ref)
assert ref.__file__.startswith("/root/reference")


def main():
    model, ids = H.tiny_model_and_ids()
    out = {"model": H.MODEL_KW, "tokens": H.TOKENS, "cases": {}}
    base = evaluate_with_compression(model, input_ids=ids, compress_fn=None, show_progress=False, return_nlls=True)
    out["cases"]["baseline"] = {"nlls": base["nlls"], "lengths": base["cache_lengths"], "perplexity": base["perplexity"]}
    for name, method, kwargs in H.CASES:
        fn = ref.get_compress_fn(method)
        r = evaluate_with_compression(model, input_ids=ids, compress_fn=fn, compress_kwargs=kwargs,
                                      skip_layers=H.SKIP, show_progress=False, return_nlls=True)
        out["cases"][name] = {"nlls": r["nlls"], "lengths": r["cache_lengths"], "perplexity": r["perplexity"],
                              "final_cache_size": r["final_cache_size"]}
        print(name, r["cache_lengths"], round(r["perplexity"], 3))
    # the attention-score harness: OUR loop (kvcompress/evaluate_attention.py) driven with the REAL reference's
    # h2o_attention_compress and H2OAttentionManager (eager attention: the model must return its weights)
    import kvcompress.evaluate_attention as EA
    from kvcompress_ref.methods.h2o_attention import H2OAttentionManager as RefManager
    from kvcompress_ref.methods.h2o_attention import h2o_attention_compress as ref_h2o_attention

    eager_model, ids = H.tiny_model_and_ids(attn_implementation="eager")
    EA.h2o_attention_compress = ref_h2o_attention
    manager = RefManager(num_layers=H.MODEL_KW["num_hidden_layers"], num_heads=H.MODEL_KW["num_attention_heads"],
                         device=torch.device("cpu"), **H.ATTN_KW)
    # skip_layers=[]: with transformers 5 the eager attention mask is sized from layer 0, so every layer must keep the
    # same number of rows (the reference's own default [0, 1] fails inside HF there)
    r = EA.evaluate_with_attention_compression(eager_model, input_ids=ids, h2o_manager=manager, skip_layers=[],
                                               show_progress=False, return_nlls=True, **H.ATTN_KW)
    out["cases"]["h2o_attention"] = {"nlls": r["nlls"], "lengths": r["cache_lengths"], "perplexity": r["perplexity"],
                                     "final_cache_size": r["final_cache_size"]}
    print("h2o_attention", r["cache_lengths"], round(r["perplexity"], 3))
    with open(os.path.join(HERE, "harness_golden.json"), "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()
