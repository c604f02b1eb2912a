#!/usr/bin/env python
"""Board power, SM clock and throttle reasons (NVML, sampled from a thread every 5 ms) while one call runs back to back
for a few seconds: the fused vote at c4, the parity-mode snapkv_lite at c4, fix_size_l2 at c2.  Answers one question:
does the vote kernel run into the board's power limit?  Prints one JSON object."""
import json
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import lab_util  # noqa: E402

lab_util.use_lab_library_if_asked()

import pynvml as nv  # noqa: E402
import torch  # noqa: E402

import kvcompress  # noqa: E402


class Sampler(threading.Thread):
    def __init__(self, h):
        super().__init__(daemon=True)
        self.h, self.stop, self.rows = h, False, []

    def run(self):
        while not self.stop:
            try:
                self.rows.append((time.perf_counter(), nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0,
                                  nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM),
                                  nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)))
            except Exception:
                pass
            time.sleep(0.005)


def measure(h, fn, seconds):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s = Sampler(h)
    s.start()
    t0 = time.perf_counter()
    n = 0
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    while time.perf_counter() - t0 < seconds:
        for _ in range(8):
            fn()
        n += 8
        torch.cuda.synchronize()
    b.record()
    torch.cuda.synchronize()
    s.stop = True
    s.join()
    rows = [r for r in s.rows if r[0] - t0 > 0.3]   # skip the ramp
    pw = sorted(r[1] for r in rows)
    ck = sorted(r[2] for r in rows)
    reasons = 0
    for r in rows:
        reasons |= r[3]
    return {"calls": n, "ms_per_call": round(a.elapsed_time(b) / n, 3), "samples": len(rows),
            "power_w_median": pw[len(pw) // 2] if pw else None, "power_w_max": pw[-1] if pw else None,
            "sm_mhz_median": ck[len(ck) // 2] if ck else None, "sm_mhz_min": ck[0] if ck else None,
            "reasons_mask": hex(reasons)}


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 3.0
    nv.nvmlInit()
    h = nv.nvmlDeviceGetHandleByIndex(0)
    dev = torch.device("cuda", 0)
    out = {"power_limit_w": nv.nvmlDeviceGetEnforcedPowerLimit(h) / 1000.0,
           "power_limit_default_w": nv.nvmlDeviceGetPowerManagementDefaultLimit(h) / 1000.0,
           "sm_max_mhz": nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)}
    L, B, H, G, S, D = 32, 16, 8, 4, 32768, 128
    kv = []
    for layer in range(L):
        g = torch.Generator(device=dev).manual_seed(layer)
        kv.append((torch.randn(B, H, S, D, generator=g, device=dev, dtype=torch.bfloat16),
                   torch.randn(B, H, S, D, generator=g, device=dev, dtype=torch.bfloat16)))
    qs = [(1.5 * torch.randn(B, H * G, 32, D, device=dev)).bfloat16() for _ in range(L)]
    lse = [torch.full((B, H * G, 32), 12.0, device=dev) for _ in range(L)]
    out["c4_vote"] = measure(h, lambda: kvcompress.snapkv_lite_compress(kv, observation_window=32, keep_size=512, obs_queries=qs), seconds)
    out["c4_vote_lse"] = measure(h, lambda: kvcompress.snapkv_lite_compress(kv, observation_window=32, keep_size=512, obs_queries=qs, obs_lse=lse), seconds)
    out["c4_snapkv_parity_mode"] = measure(h, lambda: kvcompress.snapkv_lite_compress(kv, observation_window=32, keep_size=512), seconds)
    del kv, qs, lse
    torch.cuda.empty_cache()
    L, B, H, S, D = 32, 32, 32, 4096, 80
    kv = [(torch.randn(B, H, S, D, device=dev, dtype=torch.bfloat16), torch.randn(B, H, S, D, device=dev, dtype=torch.bfloat16))
          for _ in range(L)]
    out["c2_fix_size"] = measure(h, lambda: kvcompress.fix_size_l2_compress(kv, fix_kv_size=512, keep_ratio=0.2), seconds)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
