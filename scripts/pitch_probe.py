#!/usr/bin/env python
"""Does the (batch, head) pitch of the caller's [B,H,S,D] tensors matter beyond streaming_llm?  Times the BASELINE
calls on caches whose units are views of [B,H,S+pad,D] allocations (pad = 0: the reference layout).

    python scripts/pitch_probe.py [out.json]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import lab_util  # noqa: E402,F401

import torch  # noqa: E402

import kvcompress  # noqa: E402

CASES = [
    ("c5 pyramid_kv", 32, 8, 8, 32768, 128, "pyramid_kv", dict(base_size=512)),
    ("c5 adaptive_l2", 32, 8, 8, 32768, 128, "adaptive_l2", dict(target_size=512)),
    ("c4 snapkv_lite", 32, 16, 8, 32768, 128, "snapkv_lite", dict(observation_window=32, keep_size=512)),
    ("c3 h2o_l2", 32, 32, 32, 8192, 80, "h2o_l2", dict(start_size=4, heavy_hitter_size=64, recent_size=444)),
    ("c2 fix_size_l2", 32, 32, 32, 4096, 80, "fix_size_l2", dict(fix_kv_size=512, keep_ratio=0.2)),
    ("c2 streaming_llm", 32, 32, 32, 4096, 80, "streaming_llm", dict(start_size=4, recent_size=508)),
]


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else ""
    dev = torch.device("cuda", 0)
    res = {}
    for name, L, B, H, S, D, method, kw in CASES:
        fn = kvcompress.get_compress_fn(method)
        for pad in (0, 8):
            kv = []
            for layer in range(L):
                g = torch.Generator(device=dev).manual_seed(layer)
                k = torch.empty(B, H, S + pad, D, device=dev, dtype=torch.bfloat16).normal_(generator=g)[:, :, :S]
                v = torch.empty(B, H, S + pad, D, device=dev, dtype=torch.bfloat16).normal_(generator=g)[:, :, :S]
                kv.append((k, v))
            for _ in range(3):
                fn(kv, **kw)
            torch.cuda.synchronize()
            ts = []
            for _ in range(3):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(5):
                    fn(kv, **kw)
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b) / 5)
            res[f"{name} pad={pad}"] = round(min(ts) * 1e3, 1)
            print(f"{name} pad={pad}: {res[f'{name} pad={pad}']} us", flush=True)
            del kv
            torch.cuda.empty_cache()
    if out_path:
        json.dump(res, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
