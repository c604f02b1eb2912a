"""Where the slab path's per-token overhead goes: update() vs DynamicLayer.update, SDPA on slab views vs contiguous."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cs3602-llm-inference-acceleration_b200"))
import torch, kvcompress
from transformers.cache_utils import DynamicLayer
B, H, S, D, L = 1, 32, 512, 80, 32
dt = torch.bfloat16
slab = kvcompress.KVSlabCache(L, B, H, D, 700, dt)
slab.append([(torch.randn(B, H, S, D, device="cuda").to(dt), torch.randn(B, H, S, D, device="cuda").to(dt)) for _ in range(L)])
kn, vn = torch.randn(B, H, 1, D, device="cuda").to(dt), torch.randn(B, H, 1, D, device="cuda").to(dt)
def timeit(fn, n=200):
    for _ in range(10): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    host = (time.perf_counter() - t0) / n; torch.cuda.synchronize(); wall = (time.perf_counter() - t0) / n
    return host * 1e6, wall * 1e6
def slab_updates():
    for l in range(L): slab.update(kn, vn, l)
    slab.lengths = [S] * L
layers = []
for l in range(L):
    d = DynamicLayer(); d.update(torch.randn(B, H, S, D, device="cuda").to(dt), torch.randn(B, H, S, D, device="cuda").to(dt)); layers.append(d)
base = [(d.keys, d.values) for d in layers]
def hf_updates():
    for d, (k, v) in zip(layers, base):
        d.keys, d.values = k, v
        d.update(kn, vn)
print("32 x slab.update        host %.0f us, wall %.0f us" % timeit(slab_updates))
print("32 x DynamicLayer.update host %.0f us, wall %.0f us" % timeit(hf_updates))
q = torch.randn(B, H, 1, D, device="cuda").to(dt)
ks, vs = slab[0]
ks2, vs2 = slab.k[0][:, :, :S + 1], slab.v[0][:, :, :S + 1]
kc, vc = ks2.contiguous(), vs2.contiguous()
f = torch.nn.functional.scaled_dot_product_attention
print("sdpa on slab view        host %.1f us, wall %.1f us" % timeit(lambda: f(q, ks2, vs2)))
print("sdpa on contiguous       host %.1f us, wall %.1f us" % timeit(lambda: f(q, kc, vc)))
