#!/usr/bin/env python
"""LAB: launch-plan variants for the HOST-resident (pinned slab) calls — is 39 GB/s each way on the pure copy the link
or the plan?  KVC_LAB_LIBRARY=1 python scripts/host_plan_sweep.py [out.json]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import lab_util  # noqa: E402

lab_util.use_lab_library_if_asked()

import torch  # noqa: E402

import kvcompress  # noqa: E402
from kvcompress import KVSlabCache  # noqa: E402


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else ""
    L, B, H, S, D = 32, 8, 32, 4096, 80
    dev = torch.device("cuda", 0)
    kv = []
    for layer in range(L):
        g = torch.Generator(device=dev).manual_seed(layer)
        kv.append((torch.randn(B, H, S, D, generator=g, device=dev).bfloat16(), torch.randn(B, H, S, D, generator=g, device=dev).bfloat16()))
    slab = KVSlabCache.from_legacy_cache(kv, capacity=S, pinned=True)
    del kv
    calls = [("streaming_llm", dict(start_size=4, recent_size=508), 1342177280),
             ("fix_size_l2", dict(fix_kv_size=512, keep_ratio=0.2), 1258291200)]
    variants = [{}, {"KVC_TMA_NSW": "4"}, {"KVC_TMA_NSW": "2"}, {"KVC_TMA_CTAS": "2"}, {"KVC_TMA_CTAS": "1", "KVC_TMA_NT": "256"},
                {"KVC_TMA_NT": "512"}, {"KVC_TMA_CTAS": "1", "KVC_TMA_NT": "256", "KVC_TMA_NSW": "2"}]
    res = {}
    for name, kw, nbytes in calls:
        fn = kvcompress.get_compress_fn(name)
        for var in variants:
            for k in ("KVC_TMA_CTAS", "KVC_TMA_NT", "KVC_TMA_NSW"):
                os.environ.pop(k, None)
            os.environ.update(var)
            fn(slab, **kw)
            ts = []
            for _ in range(3):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                fn(slab, **kw)
                ts.append(time.perf_counter() - t0)
            ms = min(ts) * 1e3
            key = f"{name} {var or 'default'}"
            res[key] = {"ms": round(ms, 2), "gbs_each_way": round(nbytes / ms / 1e6, 1)}
            print(key, res[key], flush=True)
    if out_path:
        json.dump(res, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
