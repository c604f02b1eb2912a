"""Run one vote launch (env-selected slicing) and report time / failure: python scripts/vote_probe.py B H G W S D L"""
import sys, time, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "cs3602-llm-inference-acceleration_b200"))
import torch
from kvcompress import _engine

B, H, G, W, S, D, L = map(int, sys.argv[1:8])
keys = [torch.randn(B, H, S, D, device="cuda").bfloat16() for _ in range(L)]
qs = [torch.randn(B, H * G, W, D, device="cuda").bfloat16() for _ in range(L)]
torch.cuda.synchronize()
t0 = time.time()
try:
    for _ in range(3):
        v = _engine.snapkv_votes(list(zip(keys, qs)), W)
    torch.cuda.synchronize()
    print("PROBE ok", sys.argv[1:8], os.environ.get("KVC_VOTE_TS"), os.environ.get("KVC_VOTE_PEND"), f"{(time.time() - t0) * 1e3 / 3:.2f} ms/launch", float(v[0].float().sum()))
except Exception as e:
    print("PROBE FAIL", sys.argv[1:8], os.environ.get("KVC_VOTE_TS"), os.environ.get("KVC_VOTE_PEND"), f"after {time.time() - t0:.1f} s", str(e).splitlines()[0])
