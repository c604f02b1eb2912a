#!/usr/bin/env python
"""Time the fused vote call (snapkv_lite_compress(obs_queries[, obs_lse])) on the c4 / c2 vote shapes with the library
in the tree; prints ms and GB/s per shape plus the SM clock.  Used to A/B kernel variants across builds."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import lab_util  # noqa: E402

lab_util.use_lab_library_if_asked()   # KVC_AB_LIBRARY=<file in csrc/>: time another build side by side

import torch  # noqa: E402

import kvcompress  # noqa: E402


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "run"
    out_path = sys.argv[2] if len(sys.argv) > 2 else ""
    dev = torch.device("cuda", 0)
    res = {}
    for name, L, B, H, G, S, D in [("c4_vote", 32, 16, 8, 4, 32768, 128), ("c2_vote", 32, 32, 32, 1, 4096, 80)]:
        kv = []
        for layer in range(L):
            g = torch.Generator(device=dev).manual_seed(layer)
            kv.append((torch.randn(B, H, S, D, generator=g, device=dev, dtype=torch.bfloat16),
                       torch.randn(B, H, S, D, generator=g, device=dev, dtype=torch.bfloat16)))
        qs = [(1.5 * torch.randn(B, H * G, 32, D, device=dev)).bfloat16() for _ in range(L)]
        lse = [torch.full((B, H * G, 32), 12.0, device=dev) for _ in range(L)]   # any finite value times the same
        for mode, kw in (("two_pass", {}), ("lse", {"obs_lse": lse})):
            nbytes = 2 * B * H * D * L * ((2 if mode == "two_pass" else 1) * (S - 32) + 4 * 512)
            fn = lambda: kvcompress.snapkv_lite_compress(kv, observation_window=32, keep_size=512, obs_queries=qs, **kw)
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(3):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(4):
                    fn()
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b) / 4)
            ms = min(ts)
            res[f"{name} {mode}"] = {"ms": round(ms, 3), "gbs": round(nbytes / ms / 1e6, 1)}
            print(tag, name, mode, res[f"{name} {mode}"], flush=True)
        del kv, qs, lse
        torch.cuda.empty_cache()
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(0)
        res["sm_mhz_after"] = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
    except Exception:
        pass
    if out_path:
        prev = json.load(open(out_path)) if os.path.exists(out_path) else {}
        prev[tag] = res
        json.dump(prev, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
