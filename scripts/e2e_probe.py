#!/usr/bin/env python
"""Where does the host-resident (pinned slab, stored norms) step spend its time?  Per call: wall clock of the public
function, device time of its kernel (CUDA events), bytes each way; then the same 8 calls of a c2 step queued back to
back with ONE synchronise at the end.

    python scripts/e2e_probe.py [--slab 8] [--out gpurun_out/e2e_probe.json]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import lab_util  # noqa: E402,F401

import torch  # noqa: E402

import kvcompress  # noqa: E402
from kvcompress import KVSlabCache, _engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--slab", type=int, default=8)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    L, H, S, D, B = 32, 32, 4096, 80, args.slab
    dev = torch.device("cuda", 0)
    kv = []
    for layer in range(L):
        g = torch.Generator(device=dev).manual_seed(layer)
        k = torch.randn(B, H, S, D, generator=g, device=dev) * torch.exp(0.35 * torch.randn(B, H, S, 1, generator=g, device=dev))
        kv.append((k.bfloat16(), torch.randn(B, H, S, D, generator=g, device=dev).bfloat16()))
    t0 = time.perf_counter()
    slab = KVSlabCache.from_legacy_cache(kv, capacity=S, pinned=True)
    build_s = time.perf_counter() - t0
    del kv
    calls = [("streaming_llm", dict(start_size=4, recent_size=508)),
             ("fix_size_l2", dict(fix_kv_size=512, keep_ratio=0.2, strategy="keep_low"))]
    res = {"slab_streams": B, "pinned_slab_build_s": round(build_s, 3), "calls": {}}
    for name, kw in calls:
        fn = kvcompress.get_compress_fn(name)
        for _ in range(2):
            fn(slab, **kw)
        walls, devs = [], []
        for _ in range(4):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            a.record()
            out = fn(slab, **kw)
            b.record()
            torch.cuda.synchronize()
            walls.append((time.perf_counter() - t0) * 1e3)
            devs.append(a.elapsed_time(b))
        nbytes = sum(k.numel() * 2 + v.numel() * 2 for (k, v), (k0, _) in zip(out, slab) if k.data_ptr() != k0.data_ptr())
        res["calls"][name] = {"wall_ms": round(min(walls), 2), "events_ms": round(min(devs), 2), "bytes_each_way": nbytes,
                              "gbs_each_way_at_wall": round(nbytes / min(walls) / 1e6, 1)}
        print(name, res["calls"][name], flush=True)
        del out
    # non-blocking: 8 calls queued back to back, one synchronise
    if "non_blocking" in kvcompress.streaming_llm_compress.__doc__:
        def step():
            outs = []
            for _ in range(4):
                for name, kw in calls:
                    outs.append(kvcompress.get_compress_fn(name)(slab, non_blocking=True, **kw))
            torch.cuda.synchronize()
            return outs
        step()
        t0 = time.perf_counter()
        for _ in range(3):
            step()
        res["queued_step_ms"] = round((time.perf_counter() - t0) / 3 * 1e3, 2)
        print("queued step", res["queued_step_ms"], flush=True)

    # two streams: the pure-copy call and the scattered-row call in flight together (do they fill each other's gaps on
    # the link, or is the link simply full?)
    if "non_blocking" in kvcompress.streaming_llm_compress.__doc__:
        s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

        def two_stream_step():
            outs = []
            for _ in range(4):
                with torch.cuda.stream(s1):
                    outs.append(kvcompress.get_compress_fn(calls[0][0])(slab, non_blocking=True, **calls[0][1]))
                with torch.cuda.stream(s2):
                    outs.append(kvcompress.get_compress_fn(calls[1][0])(slab, non_blocking=True, **calls[1][1]))
            torch.cuda.synchronize()
            return outs
        two_stream_step()
        t0 = time.perf_counter()
        for _ in range(3):
            two_stream_step()
        res["two_stream_step_ms"] = round((time.perf_counter() - t0) / 3 * 1e3, 2)
        print("two-stream step", res["two_stream_step_ms"], flush=True)

    def blocking_step():
        for _ in range(4):
            for name, kw in calls:
                kvcompress.get_compress_fn(name)(slab, **kw)
    blocking_step()
    t0 = time.perf_counter()
    for _ in range(3):
        blocking_step()
    res["blocking_step_ms"] = round((time.perf_counter() - t0) / 3 * 1e3, 2)
    print("blocking step", res["blocking_step_ms"], flush=True)
    if args.out:
        json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
