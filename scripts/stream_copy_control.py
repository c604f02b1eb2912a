#!/usr/bin/env python
"""is streaming_llm's 0.915 of the copy peak at c2 the kernel or the address set?

Moves EXACTLY the bytes streaming_llm_compress(4, 508) moves at BASELINE c2 — 32 layers x 1024 (batch, head) units x
{K, V} x rows [0,4) and [3588,4096) of a [32,32,4096,80] bf16 tensor into a dense [32,32,512,80] tensor — four ways:

  product     kvcompress.streaming_llm_compress (kvc_fused_tma_kernel, one launch for all layers)
  order1/2    (round-2 lab runs only: the kernel with other CTA -> unit mappings, KVC_TMA_ORDER — identical results,
              the knob has been removed; profiles/r02_stream_copy_control.json keeps the numbers)
  ldg         a plain LDG.128 / STG.128 kernel (scripts/lab/copyctl.cu), one launch per tensor, 1 / 4 CTAs per unit
  memcpy2d    cudaMemcpy2DAsync, two calls per tensor (sink rows, tail rows)
  contiguous  the same number of bytes as one dense copy (torch copy_): the copy-peak reference on this box

    python scripts/stream_copy_control.py [--seq-len 4096] [--out gpurun_out/stream_control.json]
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import lab_util  # noqa: E402

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seq-len", type=int, default=4096)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--only", default="")
    ap.add_argument("--pad-rows", type=int, default=0,
                    help="allocate [B,H,S+pad,D] and work on the [:, :, :S] view: a unit pitch that is not 640 KB")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    _engine = lab_util.use_lab_library_if_asked()
    import kvcompress

    L, B, H, S, D = 32, args.batch, 32, args.seq_len, 80
    sink, tail = 4, 508
    dev = torch.device("cuda", 0)
    pad = args.pad_rows
    kv = [(torch.randn(B, H, S + pad, D, device=dev, dtype=torch.bfloat16)[:, :, :S],
           torch.randn(B, H, S + pad, D, device=dev, dtype=torch.bfloat16)[:, :, :S]) for _ in range(L)]
    outs = [(torch.empty(B, H, sink + tail, D, device=dev, dtype=torch.bfloat16),
             torch.empty(B, H, sink + tail, D, device=dev, dtype=torch.bfloat16)) for _ in range(L)]
    nbytes = 2 * L * B * H * (sink + tail) * D * 2 * 2  # read + write, K and V
    lab = os.path.join(lab_util.ROOT, "scripts", "lab", "libcopyctl.so")
    if not os.path.exists(lab):
        subprocess.check_call(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
                               "-shared", "-o", lab, os.path.join(lab_util.ROOT, "scripts", "lab", "copyctl.cu")])
    ctl = ctypes.CDLL(lab)
    ctl.copyctl_ldg.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    ctl.copyctl_memcpy2d.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def product():
        return kvcompress.streaming_llm_compress(kv, start_size=sink, recent_size=tail)

    def ldg(ctas):
        def run():
            for (k, v), (ko, vo) in zip(kv, outs):
                for t, o in ((k, ko), (v, vo)):
                    rc = ctl.copyctl_ldg(t.data_ptr(), o.data_ptr(), B * H, (S + pad) * D * 2, D * 2, sink, tail, S, ctas, stream)
                    assert rc == 0, rc
        return run

    def memcpy2d():
        for (k, v), (ko, vo) in zip(kv, outs):
            for t, o in ((k, ko), (v, vo)):
                rc = ctl.copyctl_memcpy2d(t.data_ptr(), o.data_ptr(), B * H, (S + pad) * D * 2, D * 2, sink, tail, S, stream)
                assert rc == 0, rc

    flat_src = torch.empty(nbytes // 4, device=dev, dtype=torch.bfloat16)
    flat_dst = torch.empty_like(flat_src)

    def contiguous():
        flat_dst.copy_(flat_src)

    legs = [("product", product), ("ldg_1cta", ldg(1)), ("ldg_4cta", ldg(4)), ("memcpy2d", memcpy2d), ("contiguous", contiguous)]
    if os.environ.get("KVC_LAB_LIBRARY"):
        legs = [(f"product_order{os.environ.get('KVC_TMA_ORDER', '0')}", product)]
    if args.only:
        legs = [l for l in legs if l[0] in args.only.split(",")]
    res = {}
    for name, fn in legs:
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / args.steps
        name = name + (f"_pad{pad}" if pad else "")
        res[name] = {"ms": round(ms, 4), "gbs": round(nbytes / ms / 1e6, 1)}
        print(name, res[name], flush=True)
    # correctness of the controls against the product kernel
    if not args.only and not os.environ.get("KVC_LAB_LIBRARY") and not pad:
        ref = product()
        ldg(4)()
        torch.cuda.synchronize()
        assert all(torch.equal(r[0], o[0]) and torch.equal(r[1], o[1]) for r, o in zip(ref, outs)), "ldg control differs"
        for ko, vo in outs:
            ko.zero_(), vo.zero_()
        memcpy2d()
        torch.cuda.synchronize()
        assert all(torch.equal(r[0], o[0]) and torch.equal(r[1], o[1]) for r, o in zip(ref, outs)), "memcpy2d control differs"
        res["controls_bit_identical"] = True
    if pad:
        res = {k: v for k, v in res.items() if k.endswith(f"_pad{pad}")}
    res["bytes_moved"] = nbytes
    res["shape"] = f"{L} layers x (B={B}, H={H}, S={S}, D={D}) bf16, rows [0,{sink}) + [{S - tail},{S})"
    if args.out:
        prev = json.load(open(args.out)) if os.path.exists(args.out) else {}
        prev.update(res)
        json.dump(prev, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
