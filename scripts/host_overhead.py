"""B=1 decode regime: per-call cost of the public API (host-bound) — wall clock over back-to-back calls."""
import cProfile, pstats, sys, time, os, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "cs3602-llm-inference-acceleration_b200"))
import torch, kvcompress
from kvcompress import _engine
dev = torch.device("cuda")
def cache(L, B, H, S, D):
    return [(torch.randn(B, H, S, D, device=dev).bfloat16(), torch.randn(B, H, S, D, device=dev).bfloat16()) for _ in range(L)]
cases = [
    ("pythia S=513", cache(32, 1, 32, 513, 80), [("streaming_llm", {}), ("fix_size_l2", dict(fix_kv_size=512, keep_ratio=0.2)),
                                              ("h2o_l2", {}), ("snapkv_lite", {}), ("pyramid_kv", {}), ("adaptive_l2", {})]),
    ("pythia S=4096", cache(32, 1, 32, 4096, 80), [("streaming_llm", {}), ("fix_size_l2", dict(fix_kv_size=512, keep_ratio=0.2)), ("h2o_l2", {})]),
    ("llama S=32768", cache(32, 1, 8, 32768, 128), [("snapkv_lite", {}), ("adaptive_l2", {})]),
]
for name, kv, calls in cases:
    for m, kw in calls:
        fn = kvcompress.get_compress_fn(m)
        for _ in range(5): fn(kv, **kw)
        torch.cuda.synchronize()
        n = 200
        t0 = time.perf_counter()
        for _ in range(n): out = fn(kv, **kw)
        t_host = (time.perf_counter() - t0) / n
        torch.cuda.synchronize()
        t_all = (time.perf_counter() - t0) / n
        # GPU-only time of one call
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); time.sleep(0.01)
        gpu = []
        for _ in range(20):
            torch.cuda.synchronize()
            # queue a long dummy so the launch is already waiting when the GPU gets there
            x = torch.empty(1 << 28, device=dev).fill_(1.0)
            a.record(); fn(kv, **kw); b.record(); torch.cuda.synchronize()
            gpu.append(a.elapsed_time(b) * 1e3)
        print(f"{name:14s} {m:14s} host/call {t_host*1e6:7.1f} us   wall/call {t_all*1e6:7.1f} us   gpu {min(gpu):7.1f} us", flush=True)
kv = cases[0][1]
fn = kvcompress.get_compress_fn("fix_size_l2")
pr = cProfile.Profile(); pr.enable()
for _ in range(300): fn(kv, fix_kv_size=512, keep_ratio=0.2)
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(14); print(s.getvalue()[:3500])
