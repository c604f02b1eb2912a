"""Generate tests/golden/kvcompress_golden.npz by running the REAL reference.

Run only in the build container (``/root/reference`` is not on the GPU box):

    python tests/golden/make_golden.py

For every case of ``cases.py`` the reference's compress function is called on CPU tensors.
The rows it kept are recovered exactly by passing, as V, a tensor whose every element is its own
token position (the reference never looks at V when selecting).  Stored per case: per-layer
output lengths, untouched / view flags and the kept rows; plus torch's own ``torch.norm`` and
the snapkv score pipeline on a few inputs, to pin the oracle's numerics.
"""

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")

import cases as C  # noqa: E402
import kvcompress as ref  # noqa: E402  (the reference package)

assert ref.__file__.startswith("/root/reference"), ref.__file__


def to_torch(a: np.ndarray, dtype: str) -> torch.Tensor:
    if dtype == "bf16":
        return torch.from_numpy(a.view(np.int16).copy()).view(torch.bfloat16)
    return torch.from_numpy(a.copy())


def run_case(case):
    layers = C.case_cache(case)
    kv = []
    for K, V in layers:
        k = to_torch(K, case["dtype"])
        pos = torch.arange(k.size(2), dtype=torch.float32).view(1, 1, -1, 1).expand(k.shape).contiguous()
        kv.append((k, pos))
    fn = ref.get_compress_fn(case["method"])
    out = fn(kv, **case["kwargs"])
    assert len(out) == len(kv)
    meta = {"lengths": [], "untouched": [], "view": []}
    rows = {}
    for li, ((k_in, v_in), (k_out, v_out)) in enumerate(zip(kv, out)):
        meta["lengths"].append(int(k_out.size(2)))
        same = k_out is k_in and v_out is v_in
        meta["untouched"].append(bool(same))
        meta["view"].append(bool((not same) and k_out._is_view()))
        if not same:
            r = v_out[..., 0].to(torch.int64)
            assert torch.equal(v_out, r.unsqueeze(-1).expand_as(v_out).to(v_out.dtype))
            # the gathered keys must be exactly the input rows
            assert torch.equal(k_out, torch.gather(k_in, 2, r.unsqueeze(-1).expand(-1, -1, -1, k_in.size(3))))
            rows[li] = r.numpy().astype(np.int32)
    return meta, rows


def numerics_pins():
    """torch's own arithmetic on fixed inputs: norm and the snapkv importance pipeline."""
    pins = {}
    for dtype in ("f32", "bf16", "f16"):
        layers = C.make_cache(99, [700], 2, 3, 80, dtype, "spread")
        k = to_torch(layers[0][0], dtype)
        n = torch.norm(k, p=2, dim=-1)
        pins[f"pin|norm|{dtype}"] = n.float().numpy()
        for kernel in (1, 4, 5):
            prefix = n[:, :, :668]
            mx = prefix.max(dim=-1, keepdim=True)[0] + 1e-6
            imp = mx - prefix
            if kernel > 1:
                flat = imp.reshape(6, 1, 668)
                pooled = torch.nn.functional.avg_pool1d(flat, kernel_size=kernel, stride=1, padding=kernel // 2)
                imp = pooled[:, :, :668].reshape(2, 3, -1)
            pins[f"pin|snapkv_scores_k{kernel}|{dtype}"] = imp.float().numpy()
    return pins


def main():
    arrays = {}
    manifest = {}
    for case in C.all_cases():
        meta, rows = run_case(case)
        manifest[case["name"]] = meta
        for li, r in rows.items():
            arrays[f"{case['name']}|L{li}"] = r
    arrays.update(numerics_pins())
    arrays["manifest"] = np.frombuffer(json.dumps(manifest).encode(), dtype=np.uint8)
    np.savez_compressed(C.GOLDEN_NPZ, **arrays)
    size = os.path.getsize(C.GOLDEN_NPZ)
    print(f"wrote {C.GOLDEN_NPZ}: {len(manifest)} cases, {len(arrays)} arrays, {size / 1e6:.2f} MB "
          f"(torch {torch.__version__}, reference kvcompress {ref.__version__})")


if __name__ == "__main__":
    main()
