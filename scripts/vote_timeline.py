#!/usr/bin/env python
"""LAB: where does a unit of the vote kernel spend its time?  Per-CTA cycle counters written by the -DKVC_LAB build
(mbarrier waits of every role, pass boundary, tail) on the c4 vote shape, averaged over CTAs; plus the whole call with the
exponentials replaced by the identity (KVC_VOTE_DEBUG=6).  KVC_LAB_LIBRARY=1 python scripts/vote_timeline.py [out.json]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import lab_util  # noqa: E402

lab_util.use_lab_library_if_asked()

import torch  # noqa: E402

import kvcompress  # noqa: E402


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else ""
    dev = torch.device("cuda", 0)
    L, B, H, G, S, D = 32, 16, 8, 4, 32768, 128
    kv = []
    for layer in range(L):
        g = torch.Generator(device=dev).manual_seed(layer)
        kv.append((torch.randn(B, H, S, D, generator=g, device=dev, dtype=torch.bfloat16),
                   torch.randn(B, H, S, D, generator=g, device=dev, dtype=torch.bfloat16)))
    qs = [(1.5 * torch.randn(B, H * G, 32, D, device=dev)).bfloat16() for _ in range(L)]
    fn = lambda: kvcompress.snapkv_lite_compress(kv, observation_window=32, keep_size=512, obs_queries=qs)
    res = {}
    for dbg in ("0", "6", "1", "2"):
        os.environ["KVC_VOTE_DEBUG"] = dbg
        os.environ.pop("KVC_VOTE_TIMELINE", None)
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(4):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 4
        buf = torch.zeros(B * H * L, 16, dtype=torch.int64, device=dev)
        os.environ["KVC_VOTE_TIMELINE"] = hex(buf.data_ptr())
        fn()
        torch.cuda.synchronize()
        os.environ.pop("KVC_VOTE_TIMELINE", None)
        t = buf.double().mean(0).tolist()
        # wall-clock view: globaltimer start / end and SM id of every CTA -> effective SM clock and gaps between CTAs
        rec = buf.cpu()
        cyc, ns = rec[:, 11].double(), (rec[:, 10] - rec[:, 9]).double()
        eff_mhz = float((cyc.sum() / ns.sum()) * 1e3)
        gaps = []
        for sm in rec[:, 8].unique().tolist():
            rows = rec[rec[:, 8] == sm]
            rows = rows[rows[:, 9].argsort()]
            gaps += ((rows[1:, 9] - rows[:-1, 10]).double() / 1e3).tolist()
        gaps = torch.tensor(gaps)
        wall = {"effective_sm_mhz": round(eff_mhz, 1), "cta_wall_us_mean": round(float(ns.mean()) / 1e3, 1),
                "gap_between_ctas_us_mean": round(float(gaps.mean()), 2), "gap_us_max": round(float(gaps.max()), 1),
                "ctas_per_sm_max": int(torch.bincount(rec[:, 8]).max())}
        names = ["math g0: wait accumulator, pass 1", "math g0: wait accumulator, pass 2", "vote phase of the unit",
                 "producer: wait free ring slot", "MMA issuer: wait accumulator drained", "MMA issuer: wait tile landed",
                 "pass boundary", "tail (pool, select, gather)"]
        mhz = 1965.0
        res[f"debug={dbg}"] = {"call_ms": round(ms, 3), "wall": wall,
                               "per_unit_us_at_1965MHz": {n: round(v / mhz, 1) for n, v in zip(names, t[:8])}}
        print(f"debug={dbg} call {ms:.3f} ms", json.dumps(wall), json.dumps(res[f"debug={dbg}"]["per_unit_us_at_1965MHz"]), flush=True)
    if out_path:
        json.dump(res, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
