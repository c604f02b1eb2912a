#!/usr/bin/env python
"""Energy view of the fused vote call at c4 (lab build): for every stage-isolation mode of KVC_VOTE_DEBUG, the burst time
(4 calls after an idle pause) and the sustained time, board power and SM clock over a few seconds of back-to-back
calls.  At the 1000 W cap a sustained call costs (ms) joules, so the differences between modes are the energy of the
stage that was switched off.  KVC_LAB_LIBRARY=1 python scripts/vote_energy.py [seconds] [out.json]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import lab_util  # noqa: E402

lab_util.use_lab_library_if_asked()

import pynvml as nv  # noqa: E402
import torch  # noqa: E402

import kvcompress  # noqa: E402
from power_probe import measure  # noqa: E402

MODES = [("0", "product"), ("7", "pass 2 re-reads 8 L2-resident tiles (no HBM traffic in pass 2)"),
         ("8", "both passes read 8 L2-resident tiles (no K traffic from HBM at all)"),
         ("6", "exp2 replaced by the identity"), ("1", "copies + MMAs, no math"), ("2", "copies only")]


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 3.0
    out_path = sys.argv[2] if len(sys.argv) > 2 else ""
    nv.nvmlInit()
    h = nv.nvmlDeviceGetHandleByIndex(0)
    dev = torch.device("cuda", 0)
    L, B, H, G, S, D = 32, 16, 8, 4, 32768, 128
    kv = []
    for layer in range(L):
        g = torch.Generator(device=dev).manual_seed(layer)
        kv.append((torch.randn(B, H, S, D, generator=g, device=dev, dtype=torch.bfloat16),
                   torch.randn(B, H, S, D, generator=g, device=dev, dtype=torch.bfloat16)))
    qs = [(1.5 * torch.randn(B, H * G, 32, D, device=dev)).bfloat16() for _ in range(L)]
    lse = [torch.full((B, H * G, 32), 12.0, device=dev) for _ in range(L)]
    out = {}
    for with_lse in (False, True):
        kw = {"obs_lse": lse} if with_lse else {}
        fn = lambda: kvcompress.snapkv_lite_compress(kv, observation_window=32, keep_size=512, obs_queries=qs, **kw)
        for mode, what in MODES:
            if with_lse and mode in ("7",):
                continue
            os.environ["KVC_VOTE_DEBUG"] = mode
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            bursts = []
            for _ in range(3):
                time.sleep(0.5)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(4):
                    fn()
                b.record()
                torch.cuda.synchronize()
                bursts.append(a.elapsed_time(b) / 4)
            rec = measure(h, fn, seconds)
            rec["burst_ms"] = round(min(bursts), 3)
            rec["what"] = what
            out[("lse " if with_lse else "two_pass ") + "debug=" + mode] = rec
            print(("lse " if with_lse else "two_pass ") + "debug=" + mode, rec, flush=True)
    os.environ["KVC_VOTE_DEBUG"] = "0"
    if out_path:
        json.dump(out, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
