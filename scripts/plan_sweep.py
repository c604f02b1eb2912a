#!/usr/bin/env python
"""LAB: launch-plan variants (threads per CTA, resident CTAs, staging slots) for the few-unit configs — c1 (fp32, B = 1,
S = 2048, 960 units) and c2 at decode steady state with B = 1 (1024 units).  KVC_LAB_LIBRARY=1 python scripts/plan_sweep.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import lab_util  # noqa: E402

lab_util.use_lab_library_if_asked()

import torch  # noqa: E402

import kvcompress  # noqa: E402


def timed(fn, reps=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # queue a long fill first so that the launches are already waiting when the GPU gets to them
        x = torch.empty(1 << 28, device="cuda").fill_(1.0)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / reps)
        del x
    return best


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else ""
    dev = torch.device("cuda", 0)
    res = {}
    g = torch.Generator(device=dev).manual_seed(0)
    c1 = [(torch.randn(1, 32, 2048, 80, generator=g, device=dev), torch.randn(1, 32, 2048, 80, generator=g, device=dev)) for _ in range(32)]
    st = [(torch.randn(1, 32, 513, 80, generator=g, device=dev).bfloat16(), torch.randn(1, 32, 513, 80, generator=g, device=dev).bfloat16())
          for _ in range(32)]
    cases = [("c1 l2_compress", lambda: kvcompress.l2_compress(c1, keep_ratio=0.8, prune_after=1000, skip_layers=[0, 1]), 2643148800),
             ("steady_b1 fix_size_l2", lambda: kvcompress.fix_size_l2_compress(st, fix_kv_size=512, keep_ratio=0.2), 377702400),
             ("steady_b1 streaming_llm", lambda: kvcompress.streaming_llm_compress(st, start_size=4, recent_size=508), 335544320)]
    variants = [{}, {"KVC_TMA_CTAS": "2"}, {"KVC_TMA_CTAS": "1", "KVC_TMA_NT": "256"}, {"KVC_TMA_NT": "512"},
                {"KVC_TMA_NSW": "4"}, {"KVC_TMA_NSW": "2"}, {"KVC_TMA_UPC": "2"}]
    for name, fn, nbytes in cases:
        for var in variants:
            for k in ("KVC_TMA_CTAS", "KVC_TMA_NT", "KVC_TMA_NSW", "KVC_TMA_UPC"):
                os.environ.pop(k, None)
            os.environ.update(var)
            ms = timed(fn)
            key = f"{name} {var or 'default'}"
            res[key] = {"us": round(ms * 1e3, 1), "gbs": round(nbytes / ms / 1e6, 1)}
            print(key, res[key], flush=True)
    if out_path:
        json.dump(res, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
