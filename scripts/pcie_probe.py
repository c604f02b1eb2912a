"""PCIe ceilings on this box next to what the zero-copy compress path gets: python scripts/pcie_probe.py"""
import os, sys, time, json
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "cs3602-llm-inference-acceleration_b200"))
import torch
import kvcompress
try:
    import bench
    bench.bind_to_gpu_cpus(0)
except Exception as e:  # noqa: BLE001
    print("bind failed", e)

dev = torch.device("cuda:0")
N = 4 << 30
h = torch.empty(N, dtype=torch.uint8).pin_memory()
h2 = torch.empty(N, dtype=torch.uint8).pin_memory()
d = torch.empty(N, dtype=torch.uint8, device=dev)
d2 = torch.empty(N, dtype=torch.uint8, device=dev)
h.fill_(1); h2.fill_(2)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
out = {}

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps

out["memcpy_h2d_GBs"] = N / timed(lambda: d.copy_(h, non_blocking=True)) / 1e9
out["memcpy_d2h_GBs"] = N / timed(lambda: h2.copy_(d2, non_blocking=True)) / 1e9
def duplex():
    with torch.cuda.stream(s1):
        d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)
t = timed(duplex)
out["memcpy_duplex_each_GBs"] = N / t / 1e9
del d, d2, h, h2

L, B, H, S, D = 30, 8, 32, 4096, 80
kv = [(torch.randn(B, H, S, D).bfloat16().pin_memory(), torch.randn(B, H, S, D).bfloat16().pin_memory()) for _ in range(L)]
e = 2
def run(name, fn, h2d_rows, d2h_rows):
    t = timed(fn, reps=2)
    out[name] = {"ms": t * 1e3, "h2d_GBs": L * B * H * h2d_rows * D * e / t / 1e9, "d2h_GBs": L * B * H * d2h_rows * D * e / t / 1e9}
# scan-dominated: fix_size_l2 reads R = 3994 key rows + 512 K rows again + 512 V rows, writes 2 x 512 rows
run("zero_copy_fix_size", lambda: kvcompress.fix_size_l2_compress(kv, fix_kv_size=512, keep_ratio=0.2, skip_layers=[]), 3994 + 410 + 512 + 102, 1024)
# pure copy both ways: streaming_llm keeping 4 + 3000 rows of K and V
run("zero_copy_streaming_3004", lambda: kvcompress.streaming_llm_compress(kv, start_size=4, recent_size=3000), 2 * 3004, 2 * 3004)
run("zero_copy_streaming_512", lambda: kvcompress.streaming_llm_compress(kv, start_size=4, recent_size=508), 2 * 512, 2 * 512)
print("PCIE", json.dumps(out))
