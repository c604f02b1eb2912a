#!/usr/bin/env python
"""Does the page size of the pinned host cache matter for the zero-copy compress path?  The same fix_size_l2 /
streaming_llm calls on a host-resident cache whose K, V and norms live (a) in torch's pinned allocations
(cudaHostAlloc) and (b) in 2 MB-aligned anonymous memory advised MADV_HUGEPAGE, touched, then cudaHostRegister'ed.
Prints THP status, whether the kernel backed the region with huge pages (AnonHugePages in smaps), and the call times."""
import ctypes
import json
import mmap
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import lab_util  # noqa: E402,F401

import numpy as np  # noqa: E402
import torch  # noqa: E402

import kvcompress  # noqa: E402
from kvcompress import KVSlabCache  # noqa: E402

MADV_HUGEPAGE = 14
libc = ctypes.CDLL("libc.so.6", use_errno=True)


class Registered:
    """2 MB-aligned anonymous mapping, MADV_HUGEPAGE, touched, registered with CUDA; hands out tensors."""

    def __init__(self, nbytes):
        self.nbytes = (nbytes + (2 << 20) - 1) & ~((2 << 20) - 1)
        self.mm = mmap.mmap(-1, self.nbytes + (2 << 20), flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
        base = ctypes.addressof(ctypes.c_char.from_buffer(self.mm))
        self.ptr = (base + (2 << 20) - 1) & ~((2 << 20) - 1)
        rc = libc.madvise(ctypes.c_void_p(self.ptr), ctypes.c_size_t(self.nbytes), MADV_HUGEPAGE)
        self.madvise_rc = rc
        arr = (ctypes.c_char * self.nbytes).from_address(self.ptr)
        self.np = np.frombuffer(arr, dtype=np.uint8)
        self.np[:] = 0                                   # touch: fault the pages in
        err = torch.cuda.cudart().cudaHostRegister(self.ptr, self.nbytes, 0)
        self.reg_err = int(err) if not isinstance(err, int) else err
        self.off = 0

    def tensor(self, shape, dtype):
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        off = (self.off + 255) & ~255
        t = torch.from_numpy(self.np[off:off + n]).view(dtype).view(shape)
        self.off = off + n
        return t


class HostCache:
    """Minimal container the compress functions accept: (K, V) views + stored norms."""

    def __init__(self, kv, norms):
        self.kv, self.norms = kv, norms

    def to_legacy_cache(self):
        return self.kv

    def key_norm_layers(self):
        return self.norms


def anon_huge_kb():
    tot = 0
    for line in open("/proc/self/smaps"):
        if line.startswith("AnonHugePages:"):
            tot += int(line.split()[1])
    return tot


def main():
    out = {"thp_enabled": open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip(),
           "thp_defrag": open("/sys/kernel/mm/transparent_hugepage/defrag").read().strip()}
    L, B, H, S, D = 32, 8, 32, 4096, 80
    dev = torch.device("cuda", 0)
    kv = []
    for layer in range(L):
        g = torch.Generator(device=dev).manual_seed(layer)
        k = torch.randn(B, H, S, D, generator=g, device=dev) * torch.exp(0.35 * torch.randn(B, H, S, 1, generator=g, device=dev))
        kv.append((k.bfloat16(), torch.randn(B, H, S, D, generator=g, device=dev).bfloat16()))
    slab = KVSlabCache.from_legacy_cache(kv, capacity=S, pinned=True)      # (a) torch's pinned allocator
    nbytes = sum(k.numel() * 2 * 2 + k.numel() // D * 2 for k, _ in kv) + (64 << 20)
    before = anon_huge_kb()
    reg = Registered(nbytes)
    out["madvise_rc"], out["cudaHostRegister_rc"] = reg.madvise_rc, reg.reg_err
    out["anon_huge_pages_mb"] = (anon_huge_kb() - before) // 1024
    out["region_mb"] = reg.nbytes >> 20
    hk, hn = [], []
    for l in range(L):                                                       # (b) huge-page backed, registered
        k, v = reg.tensor((B, H, S, D), torch.bfloat16), reg.tensor((B, H, S, D), torch.bfloat16)
        n = reg.tensor((B, H, S), torch.bfloat16)
        k.copy_(slab[l][0]); v.copy_(slab[l][1]); n.copy_(slab.key_norms(l))
        hk.append((k, v)); hn.append(n)
    out["is_pinned"] = bool(hk[0][0].is_pinned())
    host = HostCache(hk, hn)
    calls = [("streaming_llm", dict(start_size=4, recent_size=508)),
             ("fix_size_l2", dict(fix_kv_size=512, keep_ratio=0.2, strategy="keep_low"))]
    for tag, cache in (("torch_pinned", slab), ("hugepage_registered", host), ("torch_pinned_again", slab)):
        for name, kw in calls:
            fn = kvcompress.get_compress_fn(name)
            for _ in range(2):
                res = fn(cache, **kw)
            ts = []
            for _ in range(4):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                res = fn(cache, **kw)
                torch.cuda.synchronize()
                ts.append((time.perf_counter() - t0) * 1e3)
            out[f"{tag} {name} ms"] = round(min(ts), 2)
            print(tag, name, out[f"{tag} {name} ms"], flush=True)
    # same rows out?
    a = kvcompress.fix_size_l2_compress(slab, **calls[1][1])
    b = kvcompress.fix_size_l2_compress(host, **calls[1][1])
    out["same_output"] = all(torch.equal(x[0], y[0]) and torch.equal(x[1], y[1]) for x, y in zip(a, b))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
