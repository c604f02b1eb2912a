"""Prefill-sized append into the slab (copy + norms): GB/s of 2*e*D*T*B*H*L read + the same written."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cs3602-llm-inference-acceleration_b200"))
import torch, kvcompress
for (L, B, H, T, D) in [(8, 32, 32, 4096, 80), (8, 16, 8, 32768, 128), (32, 32, 32, 1, 80), (32, 32, 32, 16, 80)]:
    kv = [(torch.randn(B, H, T, D, device="cuda").bfloat16(), torch.randn(B, H, T, D, device="cuda").bfloat16()) for _ in range(L)]
    slab = kvcompress.KVSlabCache(L, B, H, D, T + 8, torch.bfloat16)
    for _ in range(2):
        slab.lengths = [0] * L
        slab.append(kv)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    a.record()
    for _ in range(n):
        slab.lengths = [0] * L
        slab.append(kv)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / n
    nbytes = 2 * 2 * L * B * H * T * D * 2
    print(f"append L{L} B{B} H{H} T{T} D{D}: {ms*1e3:.0f} us, {nbytes/ms/1e6:.0f} GB/s (read+write)")
    # norms vs torch
    want = torch.linalg.vector_norm(kv[0][0].float(), dim=-1).to(torch.bfloat16)
    print("   norms equal fraction", float((slab.key_norms(0) == want).float().mean()))
    del kv, slab
