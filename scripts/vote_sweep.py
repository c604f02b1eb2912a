#!/usr/bin/env python
"""LAB: sweep the vote kernel's L2 prefetch distance (KVC_VOTE_PF) and stage isolation (KVC_VOTE_DEBUG) on the c4 / c2
vote shapes.  Needs the lab library (scripts/build_lab.sh); run as  KVC_LAB_LIBRARY=1 python scripts/vote_sweep.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import lab_util  # noqa: E402

import torch  # noqa: E402


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else ""
    _engine = lab_util.use_lab_library_if_asked()
    import kvcompress

    dev = torch.device("cuda", 0)
    res = {}
    shapes = [("c4_vote", 32, 16, 8, 4, 32768, 128), ("c2_vote", 32, 32, 32, 1, 4096, 80)]
    for name, L, B, H, G, S, D in shapes:
        kv = []
        for layer in range(L):
            g = torch.Generator(device=dev).manual_seed(layer)
            kv.append((torch.randn(B, H, S, D, generator=g, device=dev, dtype=torch.bfloat16),
                       torch.randn(B, H, S, D, generator=g, device=dev, dtype=torch.bfloat16)))
        qs = [(1.5 * torch.randn(B, H * G, 32, D, device=dev)).bfloat16() for _ in range(L)]
        nbytes = 2 * B * H * D * L * (2 * (S - 32) + 4 * 512)
        combos = [("0", pf, "0") for pf in ("0", "2", "4", "8", "16", "32")]
        combos += [(dbg, pf, "0") for dbg in ("1", "2") for pf in ("0", "8")]
        combos += [("0", pf, poly) for poly in ("1", "2") for pf in ("0", "8")]
        for dbg, pf, poly in combos:
            if True:
                os.environ["KVC_VOTE_PF"], os.environ["KVC_VOTE_DEBUG"], os.environ["KVC_VOTE_POLY"] = pf, dbg, poly
                fn = lambda: kvcompress.snapkv_lite_compress(kv, observation_window=32, keep_size=512, obs_queries=qs)
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
                key = f"{name} debug={dbg} pf={pf} poly={poly}"
                res[key] = {"ms": round(ms, 3), "gbs": round(nbytes / ms / 1e6, 1)}
                print(key, res[key], flush=True)
        del kv, qs
        torch.cuda.empty_cache()
    os.environ["KVC_VOTE_DEBUG"] = "0"
    if out_path:
        json.dump(res, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
