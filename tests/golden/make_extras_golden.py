"""Generate tests/golden/extras_golden.json with the REAL reference (build container only):
evict_for_space and h2o_attention_compress with an H2OAttentionManager over a short decode sequence.

    python tests/golden/make_extras_golden.py
"""

import importlib.util
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.dont_write_bytecode = True
import extras_cases as E  # noqa: E402

spec = importlib.util.spec_from_file_location("kvcompress_ref", "/root/reference/kvcompress/__init__.py",
                                              submodule_search_locations=["/root/reference/kvcompress"])
ref = importlib.util.module_from_spec(spec)
sys.modules["kvcompress_ref"] = ref
spec.loader.exec_module(ref)
from kvcompress_ref.methods.h2o_attention import H2OAttentionManager, h2o_attention_compress  # noqa: E402
from kvcompress_ref.methods.streaming_llm import evict_for_space  # noqa: E402


def rows_of(v_out):
    return v_out[0, :, :, 0].to(torch.int64).tolist()


out = {"evict": {}, "h2o": []}
for name, seq_lens, num_coming, start, recent, skip in E.EVICT_CASES:
    kv = E.evict_cache(seq_lens)
    res = evict_for_space(kv, num_coming, start_size=start, recent_size=recent, skip_layers=skip)
    out["evict"][name] = {"lengths": [k.size(2) for k, _ in res], "untouched": [r[0] is i[0] for r, i in zip(res, kv)],
                          "rows": [rows_of(v) for _, v in res]}

c = E.H2O_CASE
mgr = H2OAttentionManager(start_size=c["start_size"], heavy_hitter_size=c["heavy_hitter_size"], recent_size=c["recent_size"],
                          num_layers=c["layers"], num_heads=c["heads"], decay_factor=c["decay_factor"])
seq_len = c["S"]
for step in range(c["steps"]):
    kv, attn = E.h2o_inputs(step, seq_len)
    res = h2o_attention_compress(kv, attention_scores=attn, h2o_manager=mgr, start_size=c["start_size"],
                                 heavy_hitter_size=c["heavy_hitter_size"], recent_size=c["recent_size"],
                                 skip_layers=c["skip_layers"])
    out["h2o"].append({"seq_len": seq_len, "lengths": [k.size(2) for k, _ in res], "rows": [rows_of(v) for _, v in res]})
    seq_len = res[-1][0].size(2) + 1  # the next decode step sees the compressed cache plus one token

with open(os.path.join(HERE, "extras_golden.json"), "w") as f:
    json.dump(out, f)
print({k: v["lengths"] for k, v in out["evict"].items()}, [s["lengths"] for s in out["h2o"]])
