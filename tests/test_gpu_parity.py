"""-m gpu: the sm_100a path (through the drop-in Python API and the C ABI) against the oracle,
the committed reference golden vectors, and size-independent properties at BASELINE shapes."""

import numpy as np
import pytest
import torch

import cases
import kvcompress
from gpu_util import TORCH_DTYPE, kv_to_torch, to_numpy, to_torch
from kvcompress import _engine
from kvcompress import _planner as P
from oracle import kvc_oracle as O
from test_planner import plan_for

pytestmark = pytest.mark.gpu

ALL = cases.all_cases()


@pytest.fixture(scope="module", autouse=True)
def _needs_cuda():
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    _engine.load_library()


def gather_rows(x: torch.Tensor, rows: torch.Tensor) -> torch.Tensor:
    return torch.gather(x, 2, rows.long().unsqueeze(-1).expand(-1, -1, -1, x.size(3)))


@pytest.mark.parametrize("case", ALL, ids=[c["name"] for c in ALL])
def test_golden_case(case, golden):
    """Every golden case: lengths / aliasing as the reference, kept rows valid under the tie-aware rule
    (identical to the oracle for fp32), K/V bit-exact gathers of the input."""
    data, manifest = golden
    meta = manifest[case["name"]]
    dtype = case["dtype"]
    layers = cases.case_cache(case)
    kv = kv_to_torch(layers, dtype)
    launches0 = _engine.launch_count()
    out = kvcompress.get_compress_fn(case["method"])(kv, **case["kwargs"])
    n_launch = _engine.launch_count() - launches0
    results = O.METHODS[case["method"]](layers, dtype, **case["kwargs"])
    assert [k.size(2) for k, _ in out] == meta["lengths"]
    n_gather = 0
    for li, ((k_in, v_in), (k_out, v_out), res) in enumerate(zip(kv, out, results)):
        if meta["untouched"][li]:
            assert k_out is k_in and v_out is v_in
            continue
        if meta["view"][li]:
            assert k_out._is_view() and v_out._is_view()  # aliases the input like the reference
        else:
            n_gather += 1
            assert k_out.is_contiguous() and v_out.is_contiguous()
            assert k_out.dtype == k_in.dtype and k_out.device == k_in.device
        want = torch.from_numpy(res.rows).cuda()
        if res.mode in ("none",) or res.k_sel == 0:
            assert torch.equal(k_out, gather_rows(k_in, want)) and torch.equal(v_out, gather_rows(v_in, want))
    assert n_launch == (1 if n_gather else 0), "all gathered layers of a call must share ONE launch"

    # kept rows, through the C ABI's idx_out
    plans = plan_for(case["method"], case["seq_lens"], case["kwargs"])
    out2, idx = _engine.run_plans(kv, plans, return_indices=True)
    for li, res in enumerate(results):
        if plans[li].kind != P.GATHER:
            continue
        rows = idx[li]
        k_in, v_in = kv[li]
        # gathered K/V are bit-exact copies of the rows the kernel reports
        assert torch.equal(out2[li][0], gather_rows(k_in, rows)) and torch.equal(out2[li][1], gather_rows(v_in, rows))
        assert torch.equal(out2[li][0], out[li][0]) and torch.equal(out2[li][1], out[li][1])  # deterministic
        info = O.check_layer(layers[li][0], dtype, res, rows.cpu().numpy())
        assert info["valid"], (case["name"], li, info)
        if dtype == "f32" and case["style"] != "ties":
            assert info["identical_heads"] == info["heads"], (case["name"], li, info)
            ref_rows = data[f"{case['name']}|L{li}"].astype(np.int64)
            assert np.array_equal(rows.cpu().numpy(), ref_rows)  # == the reference's own kept rows
        else:
            # ties -> lowest index is exactly the oracle's rule; allow only rounding-boundary differences
            assert info["identical_heads"] >= 0.98 * info["heads"], (case["name"], li, info)


@pytest.mark.parametrize("dtype", ["f32", "bf16", "f16"])
@pytest.mark.parametrize("D", [16, 24, 64, 80, 96, 128, 256])
def test_key_norms(dtype, D):
    """K1: fp32 norms within 1e-6 relative of the float64 norm (north_star), rounded once to the dtype."""
    if (D * (4 if dtype == "f32" else 2)) % 16:
        pytest.skip("row not 16-byte aligned")
    layers = cases.make_cache(7 + D, [777], 2, 3, D, dtype, "spread")
    K = layers[0][0]
    k = to_torch(K, dtype)
    got = _engine.key_norms(k)
    assert got.dtype == k.dtype and got.shape == (2, 3, 777)
    exact = O.exact_norms(K, dtype)
    gotf = got.float().cpu().numpy()
    if dtype == "f32":
        assert np.max(np.abs(gotf - exact) / exact) <= 1e-6
    else:
        want = O.key_norms(K, dtype)
        assert np.mean(gotf == want) > 0.999
        lo, hi = O.key_norm_interval(K, dtype)
        assert np.all((gotf >= lo) & (gotf <= hi))
    # sub-range + strided (transposed [B,S,H,D] storage) input
    base = k.permute(0, 2, 1, 3).contiguous().permute(0, 2, 1, 3)
    assert not base.is_contiguous()
    got2 = _engine.key_norms(base, 5, 700)
    assert torch.equal(got2, got[:, :, 5:700])


@pytest.mark.parametrize("dtype", ["f32", "bf16", "f16"])
@pytest.mark.parametrize("n,k", [(1, 1), (33, 5), (1000, 1), (1000, 999), (1000, 1000), (4097, 512), (32768, 480),
                                 (50000, 3000)])
def test_select_matches_stable_sort(dtype, n, k):
    """K2: exactly the stable-sort answer, for both directions, with heavy ties and negative scores."""
    rng = np.random.default_rng(n * 31 + k)
    rows = 5
    base = rng.standard_normal((rows, n)).astype(np.float32)
    base[1] = np.round(base[1] * 2) / 2          # few distinct values -> many ties
    base[2] = 1.25                               # all equal
    base[3] = np.abs(base[3]) * 1e-3             # tiny positives
    base[4, ::7] = -0.0                          # signed zeros mixed with values
    scores_np = O.store(base, dtype)
    vals = O.to_f32(scores_np, dtype)
    t = to_torch(scores_np, dtype)
    for largest in (False, True):
        got = _engine.select(t, k, largest=largest).cpu().numpy()
        # -0.0 sorts below +0.0 in the kernel's total order; mirror that in the oracle keys
        keyed = np.where(np.signbit(vals) & (vals == 0), -1e-45, vals).astype(np.float64)
        want = O.highest_k(keyed, k) if largest else O.lowest_k(keyed, k)
        assert np.array_equal(got, want), (dtype, n, k, largest)


def test_strided_and_noncontiguous_inputs():
    """Inputs living in a larger pre-allocated cache slab (stride_h != S*D) and [B,S,H,D]-stored caches."""
    torch.manual_seed(0)
    B, H, S, D, cap = 2, 4, 900, 80, 1024
    slab_k = torch.randn(B, H, cap, D, device="cuda").to(torch.bfloat16)
    slab_v = torch.randn(B, H, cap, D, device="cuda").to(torch.bfloat16)
    kv_view = [(slab_k[:, :, :S], slab_v[:, :, :S])]
    kv_dense = [(slab_k[:, :, :S].contiguous(), slab_v[:, :, :S].contiguous())]
    for fn, kw in ((kvcompress.h2o_l2_compress, {}), (kvcompress.streaming_llm_compress, {}),
                   (kvcompress.snapkv_lite_compress, {})):
        a, b = fn(kv_view, **kw), fn(kv_dense, **kw)
        assert torch.equal(a[0][0], b[0][0]) and torch.equal(a[0][1], b[0][1])
    bshd_k = torch.randn(B, S, H, D, device="cuda").to(torch.bfloat16)
    bshd_v = torch.randn(B, S, H, D, device="cuda").to(torch.bfloat16)
    kv_t = [(bshd_k.permute(0, 2, 1, 3), bshd_v.permute(0, 2, 1, 3))]
    kv_c = [(kv_t[0][0].contiguous(), kv_t[0][1].contiguous())]
    a, b = kvcompress.h2o_l2_compress(kv_t), kvcompress.h2o_l2_compress(kv_c)
    assert torch.equal(a[0][0], b[0][0]) and torch.equal(a[0][1], b[0][1])
    # different K and V strides in the same layer
    kv_mixed = [(kv_t[0][0], kv_c[0][1])]
    c = kvcompress.h2o_l2_compress(kv_mixed)
    assert torch.equal(c[0][0], b[0][0]) and torch.equal(c[0][1], b[0][1])


def test_random_strategy_follows_torch_rng_stream():
    """fix_size_l2 strategy="random": same torch.randperm calls in the same order as the reference
    (fix_size_l2.py:118-124), so the same generator state gives the same kept rows."""
    B, H, S, D = 2, 3, 700, 64
    kv = [(torch.randn(B, H, S, D, device="cuda"), torch.randn(B, H, S, D, device="cuda")) for _ in range(3)]
    torch.manual_seed(123)
    out = kvcompress.fix_size_l2_compress(kv, fix_kv_size=256, keep_ratio=0.25, strategy="random", skip_layers=[0])
    torch.manual_seed(123)
    for li in (1, 2):
        zone_end, take = S - 64, 256 - 64
        picked = torch.stack([torch.stack([torch.randperm(zone_end, device="cuda")[:take] for _ in range(H)])
                              for _ in range(B)])
        rows = torch.cat([torch.sort(picked, dim=-1)[0], torch.arange(S - 64, S, device="cuda").expand(B, H, 64)], -1)
        assert torch.equal(out[li][0], gather_rows(kv[li][0], rows))
        assert torch.equal(out[li][1], gather_rows(kv[li][1], rows))
    assert out[0][0] is kv[0][0]


def test_more_layers_than_one_launch_holds_and_mixed_groups():
    """> KVC_MAX_LAYERS_PER_LAUNCH layers are chunked; layers of different shape/dtype are grouped."""
    S, D = 300, 32
    kv = [(torch.randn(1, 2, S, D, device="cuda"), torch.randn(1, 2, S, D, device="cuda")) for _ in range(70)]
    kv.append((torch.randn(2, 1, S, 64, device="cuda").half(), torch.randn(2, 1, S, 64, device="cuda").half()))
    n0 = _engine.launch_count()
    out = kvcompress.streaming_llm_compress(kv, start_size=4, recent_size=60)
    assert _engine.launch_count() - n0 == 3  # 64 + 6 fp32 layers, then the fp16 group
    for (k, v), (ko, vo) in zip(kv, out):
        assert torch.equal(ko, torch.cat([k[:, :, :4], k[:, :, -60:]], 2))
        assert torch.equal(vo, torch.cat([v[:, :, :4], v[:, :, -60:]], 2))


def test_runs_on_the_current_stream_without_sync():
    kv = [(torch.randn(2, 4, 2000, 80, device="cuda").bfloat16(), torch.randn(2, 4, 2000, 80, device="cuda").bfloat16())]
    want = kvcompress.h2o_l2_compress(kv)
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        got = kvcompress.h2o_l2_compress(kv)
    s.synchronize()
    assert torch.equal(got[0][0], want[0][0]) and torch.equal(got[0][1], want[0][1])


def test_errors():
    kv = [(torch.randn(1, 2, 700, 12, device="cuda").bfloat16(), torch.randn(1, 2, 700, 12, device="cuda").bfloat16())]
    with pytest.raises(ValueError, match="multiple of 16"):
        kvcompress.streaming_llm_compress(kv)
    kv = [(torch.randn(1, 2, 700, 16, device="cuda").double(), torch.randn(1, 2, 700, 16, device="cuda").double())]
    with pytest.raises(ValueError, match="not supported"):
        kvcompress.streaming_llm_compress(kv)
    kv = [(torch.randn(1, 2, 2000, 16, device="cuda"), torch.randn(1, 2, 2000, 16, device="cuda"))]
    with pytest.raises(ValueError, match="Unknown strategy: nope"):
        kvcompress.fix_size_l2_compress(kv, fix_kv_size=512, strategy="nope", skip_layers=[])


def test_any_row_width_takes_the_workspace_beyond_shared_memory():
    """Row widths without a compiled kernel run the generic-width instantiation of the SAME kernel, so they get
    the workspace path too: an fp32 region beyond the on-chip key buffer at head_dim 16 (round 1: an error)."""
    S = 70000
    k = torch.zeros(1, 1, S, 16, device="cuda")
    k[0, 0, 1000:1064] = -1e-3      # 64 rows with the only non-zero norms: never kept by keep-lowest
    v = torch.arange(S, device="cuda", dtype=torch.float32).view(1, 1, S, 1).expand(1, 1, S, 16).contiguous()
    out = kvcompress.h2o_l2_compress([(k, v)])
    kept = out[0][1][0, 0, :, 0].long()
    # all-zero norms tie: lowest indices win -> sinks 0..3, heavy hitters 4..67, the last 444 rows
    want = torch.cat([torch.arange(0, 68, device="cuda"), torch.arange(S - 444, S, device="cuda")])
    assert torch.equal(kept, want)


# ----------------------------------------------------------------------------------------------
# BASELINE.json shapes (full head_dim and context, reduced batch): size-independent properties
def spread_cache(L, B, H, S, D, dtype, seed=1234):
    """BASELINE.md §3 synthetic input, generated on device."""
    out = []
    for layer in range(L):
        g = torch.Generator(device="cuda").manual_seed(seed + layer)
        k = torch.randn(B, H, S, D, generator=g, device="cuda")
        k *= torch.exp(0.35 * torch.randn(B, H, S, 1, generator=g, device="cuda"))
        k[:, :, :4] *= 0.1
        v = torch.randn(B, H, S, D, generator=g, device="cuda")
        out.append((k.to(dtype), v.to(dtype)))
    return out


def torch_key_interval(keys: torch.Tensor, rel=1e-6):
    n = torch.linalg.vector_norm(keys.double(), dim=-1)
    lo = (n * (1 - rel)).float().to(keys.dtype).float()
    hi = (n * (1 + rel)).float().to(keys.dtype).float()
    return lo, hi


def assert_valid_lowest(keys_region: torch.Tensor, sel: torch.Tensor, largest=False):
    """Tie-aware rule on device: max(lo[selected]) <= min(hi[unselected]) per (b,h); sel ascending."""
    assert torch.all(sel[..., 1:] > sel[..., :-1])
    lo, hi = torch_key_interval(keys_region)
    mask = torch.zeros_like(lo, dtype=torch.bool).scatter_(-1, sel.long(), True)
    inf = float("inf")
    if largest:
        assert torch.all(torch.where(mask, hi, inf).amin(-1) >= torch.where(mask, -inf, lo).amax(-1))
    else:
        assert torch.all(torch.where(mask, lo, -inf).amax(-1) <= torch.where(mask, inf, hi).amin(-1))


FULL = [
    ("c2_fix_size", "fix_size_l2", dict(fix_kv_size=512, keep_ratio=0.2, strategy="keep_low"), 4, 2, 32, 4096, 80, torch.bfloat16),
    ("c2_fix_size_steady", "fix_size_l2", dict(fix_kv_size=512, keep_ratio=0.2), 4, 2, 32, 513, 80, torch.bfloat16),
    ("c3_h2o", "h2o_l2", dict(start_size=4, heavy_hitter_size=64, recent_size=444), 3, 2, 32, 8192, 80, torch.bfloat16),
    ("c5_pyramid", "pyramid_kv", dict(base_size=512), 8, 2, 8, 32768, 128, torch.bfloat16),
    ("c5_adaptive", "adaptive_l2", dict(target_size=512), 2, 2, 8, 32768, 128, torch.bfloat16),
    ("c1_l2_gpu", "l2_compress", dict(keep_ratio=0.8, prune_after=1000, skip_layers=[0, 1]), 4, 1, 32, 2048, 80, torch.float32),
    ("fp32_48k", "h2o_l2", dict(start_size=4, heavy_hitter_size=64, recent_size=444), 1, 1, 4, 49152, 64, torch.float32),
]


@pytest.mark.parametrize("name,method,kwargs,L,B,H,S,D,dtype", FULL, ids=[f[0] for f in FULL])
def test_full_shape_properties(name, method, kwargs, L, B, H, S, D, dtype):
    kv = spread_cache(L, B, H, S, D, dtype)
    fn = kvcompress.get_compress_fn(method)
    out = fn(kv, **kwargs)
    plans = plan_for(method, [S] * L, kwargs)
    out2, idx = _engine.run_plans(kv, plans, return_indices=True)
    for li, p in enumerate(plans):
        if p.kind == P.KEEP:
            assert out[li][0] is kv[li][0]
            continue
        rows = idx[li]
        k_in, v_in = kv[li]
        assert out[li][0].shape == (B, H, p.out_len, D)
        assert torch.equal(out[li][0], gather_rows(k_in, rows)) and torch.equal(out[li][1], gather_rows(v_in, rows))
        assert torch.equal(rows[..., :p.sink], torch.arange(p.sink, device="cuda", dtype=torch.int32).expand(B, H, -1))
        tail = torch.arange(S - p.tail, S, device="cuda", dtype=torch.int32).expand(B, H, -1)
        assert torch.equal(rows[..., p.sink + p.k_sel:], tail)
        sel = rows[..., p.sink:p.sink + p.k_sel] - p.sel_lo
        assert sel.min() >= 0 and sel.max() < p.sel_hi - p.sel_lo
        assert_valid_lowest(k_in[:, :, p.sel_lo:p.sel_hi], sel)
    # idempotence: the compressed cache is within budget, a second call returns the same objects
    again = fn(out, **kwargs)
    if method not in ("l2_compress", "adaptive_l2"):  # ratio / gradual-zone compression keeps shrinking by design
        assert all(a[0] is b[0] for a, b in zip(again, out))


def test_full_shape_streaming_and_snapkv():
    kv = spread_cache(3, 2, 32, 4096, 80, torch.bfloat16)
    out = kvcompress.streaming_llm_compress(kv, start_size=4, recent_size=508)
    for (k, v), (ko, vo) in zip(kv, out):
        assert torch.equal(ko, torch.cat([k[:, :, :4], k[:, :, -508:]], 2))
        assert torch.equal(vo, torch.cat([v[:, :, :4], v[:, :, -508:]], 2))
    # c4: snapkv_lite at 32K context, Llama-3-8B KV shape; scores restated with torch ops on device
    kv = spread_cache(2, 2, 8, 32768, 128, torch.bfloat16)
    plans = plan_for("snapkv_lite", [32768] * 2, dict(observation_window=32, keep_size=512))
    out, idx = _engine.run_plans(kv, plans, return_indices=True)
    for li, p in enumerate(plans):
        k_in, v_in = kv[li]
        rows = idx[li]
        assert torch.equal(out[li][0], gather_rows(k_in, rows)) and torch.equal(out[li][1], gather_rows(v_in, rows))
        pre = torch.linalg.vector_norm(k_in[:, :, :p.sel_hi].float(), dim=-1).to(k_in.dtype)
        imp = (pre.max(dim=-1, keepdim=True)[0] + 1e-6) - pre
        pooled = torch.nn.functional.avg_pool1d(imp.reshape(-1, 1, p.sel_hi), 5, 1, 2).reshape(imp.shape).float()
        sel = rows[..., :p.k_sel].long()
        mask = torch.zeros_like(pooled, dtype=torch.bool).scatter_(-1, sel, True)
        lowest_taken = torch.where(mask, pooled, float("inf")).amin(-1)
        highest_left = torch.where(mask, -float("inf"), pooled).amax(-1)
        # allow one bf16 ulp: torch's CUDA avg_pool may sum in a different order than the CPU reference
        assert torch.all(lowest_taken >= highest_left * (1 - 2 ** -7))
        assert torch.equal(rows[..., p.k_sel:], torch.arange(32768 - 32, 32768, device="cuda", dtype=torch.int32).expand(2, 8, -1))


# ----------------------------------------------------------------------------------------------
# Offloaded caches: pinned host tensors are compressed in place by the GPU (zero-copy over PCIe)
@pytest.mark.parametrize("method,kwargs", [
    ("streaming_llm", dict(start_size=4, recent_size=124)),
    ("fix_size_l2", dict(fix_kv_size=256, keep_ratio=0.2, skip_layers=[0])),
    ("fix_size_l2", dict(fix_kv_size=256, keep_ratio=0.25, strategy="random", skip_layers=[])),
    ("h2o_l2", dict(start_size=4, heavy_hitter_size=32, recent_size=92)),
    ("snapkv_lite", dict(observation_window=16, keep_size=128)),
    ("pyramid_kv", dict(base_size=128, min_size=32)),
])
def test_pinned_host_cache_matches_device_cache(method, kwargs):
    """Same bytes out whether the cache lives in HBM or in pinned host memory; host in -> pinned host out,
    results complete on return (the reference's CPU path is synchronous); pageable host tensors are refused."""
    torch.manual_seed(5)
    L, B, H, S, D = 3, 2, 4, 700, 80
    host = [(torch.randn(B, H, S, D).bfloat16().pin_memory(), torch.randn(B, H, S, D).bfloat16().pin_memory())
            for _ in range(L)]
    dev = [(k.cuda(), v.cuda()) for k, v in host]
    fn = kvcompress.get_compress_fn(method)
    torch.manual_seed(99)
    out_h = fn(host, **kwargs)
    torch.manual_seed(99)
    out_d = fn(dev, **kwargs)
    if kwargs.get("strategy") == "random":  # CPU and CUDA generators differ: check structure only
        for (kh, vh), (kd, vd) in zip(out_h, out_d):
            assert kh.shape == kd.shape and kh.device.type == "cpu"
        return
    for li, ((kh, vh), (kd, vd)) in enumerate(zip(out_h, out_d)):
        assert kh.device.type == "cpu" and vh.device.type == "cpu"
        if kd is dev[li][0]:
            assert kh is host[li][0]
            continue
        assert kh.is_pinned() and vh.is_pinned()
        assert torch.equal(kh, kd.cpu()) and torch.equal(vh, vd.cpu())
    # a [B,S,H,D]-stored pinned cache (row stride != row bytes: per-row bulk copies)
    t = [(k.permute(0, 2, 1, 3).contiguous().pin_memory().permute(0, 2, 1, 3),
          v.permute(0, 2, 1, 3).contiguous().pin_memory().permute(0, 2, 1, 3)) for k, v in host]
    out_t = fn(t, **kwargs)
    for (kt, vt), (kh, vh) in zip(out_t, out_h):
        assert torch.equal(kt, kh) and torch.equal(vt, vh)
    with pytest.raises(RuntimeError, match="no CPU path"):
        fn([(k.clone(), v.clone()) for k, v in host], **kwargs)


# ----------------------------------------------------------------------------------------------
# The golden cases use small head dims (rows of 32-64 B -> the generic-width kernel).  The same presets at the
# PRODUCT row widths (compile-time widths: 128/160/192/256/320/512-byte rows) and at a few odd ones (bf16 D=72 / 40,
# fp32 D=100: 144 / 80 / 400-byte rows, generic width) against the golden-pinned oracle.
WIDTHS = [("bf16", 64), ("bf16", 80), ("bf16", 96), ("bf16", 128), ("bf16", 256), ("f16", 80), ("f32", 32),
          ("f32", 48), ("f32", 80), ("f32", 128), ("bf16", 72), ("bf16", 40), ("f32", 100)]
WIDE_PRESETS = [
    ("streaming", "streaming_llm", dict(start_size=4, recent_size=508)),
    ("h2o", "h2o_l2", dict(start_size=4, heavy_hitter_size=64, recent_size=444, skip_layers=[1])),
    ("snapkv", "snapkv_lite", dict(observation_window=32, keep_size=512)),
    ("snapkv_pool4", "snapkv_lite", dict(observation_window=16, keep_size=256, pooling_kernel=4)),
    ("pyramid", "pyramid_kv", dict(base_size=512, layer_decay=0.9, min_size=64)),
    ("adaptive", "adaptive_l2", dict(target_size=512, soft_limit=256, hard_limit=1024)),
    ("adaptive_gradual", "adaptive_l2", dict(target_size=512, soft_limit=256, hard_limit=2048)),
    ("fix_low", "fix_size_l2", dict(fix_kv_size=512, strategy="keep_low", keep_ratio=0.2, skip_layers=[0])),
    ("fix_high", "fix_size_l2", dict(fix_kv_size=256, strategy="keep_high", keep_ratio=0.3, skip_layers=[])),
    ("l2", "l2_compress", dict(keep_ratio=0.8, prune_after=100, skip_layers=[0])),
]


@pytest.mark.parametrize("dtype,D", WIDTHS, ids=[f"{d}-D{n}" for d, n in WIDTHS])
@pytest.mark.parametrize("name,method,kwargs", WIDE_PRESETS, ids=[p[0] for p in WIDE_PRESETS])
@pytest.mark.parametrize("style", ["spread", "ties"])
def test_oracle_parity_at_product_row_widths(name, method, kwargs, dtype, D, style):
    if style == "ties" and (dtype == "f32" or method == "streaming_llm"):
        pytest.skip("tie stress is the 16-bit selection case")
    case = cases._case(f"wide/{name}/{dtype}/D{D}/{style}", method, kwargs, [1300, 1300, 700], dtype=dtype, style=style,
                       B=2, H=2, D=D, seed=D * 7 + len(name))
    layers = cases.case_cache(case)
    kv = kv_to_torch(layers, dtype)
    results = O.METHODS[method](layers, dtype, **kwargs)
    plans = plan_for(method, case["seq_lens"], kwargs)
    n0 = _engine.launch_count()
    out, idx = _engine.run_plans(kv, plans, return_indices=True)
    assert _engine.launch_count() - n0 <= 1
    api = kvcompress.get_compress_fn(method)(kv, **kwargs)
    for li, res in enumerate(results):
        k_in, v_in = kv[li]
        assert api[li][0].size(2) == len(res.rows[0, 0]) if not res.untouched else api[li][0] is k_in
        if plans[li].kind != P.GATHER:
            continue
        rows = idx[li]
        assert torch.equal(out[li][0], gather_rows(k_in, rows)) and torch.equal(out[li][1], gather_rows(v_in, rows))
        assert torch.equal(api[li][0], out[li][0]) and torch.equal(api[li][1], out[li][1])
        info = O.check_layer(layers[li][0], dtype, res, rows.cpu().numpy())
        assert info["valid"], (case["name"], li, info)
        if dtype == "f32":
            assert info["identical_heads"] == info["heads"], (case["name"], li, info)
        else:
            # validity is the rule; a head differs from the oracle only when a norm sits on a rounding boundary of
            # the 16-bit dtype (the oracle sums in a different order): rare, never systematic
            assert info["identical_heads"] >= 0.9 * info["heads"], (case["name"], li, info)


# ----------------------------------------------------------------------------------------------
# Region lengths around the block sizes of the select (32 x 8 keys per warp step) and of the pooling transform
# (128 rows per warp step, 4 per lane), for every specialised pooling kernel and the generic one.
EDGE_R = [5, 7, 8, 33, 127, 128, 129, 255, 256, 257, 511, 513, 1023, 1025, 2047, 2049, 4100]


@pytest.mark.parametrize("dtype,D", [("bf16", 128), ("f32", 80), ("f16", 64)], ids=["bf16-D128", "f32-D80", "f16-D64"])
@pytest.mark.parametrize("pk", [1, 3, 4, 5, 7, 9])
def test_snapkv_region_lengths_around_block_boundaries(pk, dtype, D):
    w = 8
    for style in ("spread", "ties"):
        if style == "ties" and dtype == "f32":
            continue
        seq_lens = [r + w for r in EDGE_R]
        kwargs = dict(observation_window=w, keep_size=w + 3, pooling_kernel=pk, skip_layers=[])
        case = cases._case(f"edge/snapkv/pk{pk}/{dtype}/{style}", "snapkv_lite", kwargs, seq_lens, dtype=dtype, style=style,
                           B=1, H=2, D=D, seed=pk * 131 + D)
        layers = cases.case_cache(case)
        kv = kv_to_torch(layers, dtype)
        results = O.METHODS["snapkv_lite"](layers, dtype, **kwargs)
        plans = plan_for("snapkv_lite", seq_lens, kwargs)
        out, idx = _engine.run_plans(kv, plans, return_indices=True)
        for li, res in enumerate(results):
            if plans[li].kind != P.GATHER:
                continue
            k_in, v_in = kv[li]
            rows = idx[li]
            assert torch.equal(out[li][0], gather_rows(k_in, rows)) and torch.equal(out[li][1], gather_rows(v_in, rows))
            info = O.check_layer(layers[li][0], dtype, res, rows.cpu().numpy())
            assert info["valid"], (case["name"], seq_lens[li], info)
            if dtype == "f32":
                assert info["identical_heads"] == info["heads"], (case["name"], seq_lens[li], info)


@pytest.mark.parametrize("dtype", ["bf16", "f32"])
def test_select_region_lengths_and_large_k(dtype):
    """Few rows kept (h2o_l2) and nearly all rows kept (l2_compress at 0.97) on every edge length, ties included."""
    for style in ("spread", "ties"):
        if style == "ties" and dtype == "f32":
            continue
        seq_lens = [r + 12 for r in EDGE_R if r > 8]
        kwargs = dict(start_size=4, heavy_hitter_size=3, recent_size=8, skip_layers=[])  # a handful of rows of each region
        case = cases._case(f"edge/h2o/{dtype}/{style}", "h2o_l2", kwargs, seq_lens, dtype=dtype, style=style,
                           B=1, H=2, D=64 if dtype == "bf16" else 32, seed=20)
        layers = cases.case_cache(case)
        kv = kv_to_torch(layers, dtype)
        results = O.METHODS["h2o_l2"](layers, dtype, **kwargs)
        plans = plan_for("h2o_l2", seq_lens, kwargs)
        out, idx = _engine.run_plans(kv, plans, return_indices=True)
        for li, res in enumerate(results):
            if plans[li].kind != P.GATHER:
                continue
            rows = idx[li]
            assert torch.equal(out[li][0], gather_rows(kv[li][0], rows))
            info = O.check_layer(layers[li][0], dtype, res, rows.cpu().numpy())
            assert info["valid"], (case["name"], seq_lens[li], info)
    # keep most of a region (l2_compress, keep_ratio 0.97): k close to R on every edge length
    seq_lens = [r for r in EDGE_R if r >= 127]
    kwargs = dict(keep_ratio=0.97, prune_after=10, skip_layers=[])
    case = cases._case(f"edge/l2/{dtype}", "l2_compress", kwargs, seq_lens, dtype=dtype, style="ties" if dtype == "bf16" else "spread",
                       B=1, H=2, D=64, seed=5)
    layers = cases.case_cache(case)
    kv = kv_to_torch(layers, dtype)
    results = O.METHODS["l2_compress"](layers, dtype, **kwargs)
    plans = plan_for("l2_compress", seq_lens, kwargs)
    out, idx = _engine.run_plans(kv, plans, return_indices=True)
    for li, res in enumerate(results):
        if plans[li].kind != P.GATHER:
            continue
        info = O.check_layer(layers[li][0], dtype, res, idx[li].cpu().numpy())
        assert info["valid"], (case["name"], seq_lens[li], info)
        assert torch.equal(out[li][0], gather_rows(kv[li][0], idx[li]))


# ----------------------------------------------------------------------------------------------
# Selections larger than shared memory: radix keys / kept indices fall back to a device workspace
BIG = [
    ("l2_08_32k_f32", "l2_compress", dict(keep_ratio=0.8, prune_after=100, skip_layers=[]), 2, 1, 2, 32768, 64, torch.float32),
    ("l2_05_64k_bf16", "l2_compress", dict(keep_ratio=0.5, prune_after=100, skip_layers=[]), 1, 1, 2, 65536, 128, torch.bfloat16),
    ("h2o_120k_bf16", "h2o_l2", dict(start_size=4, heavy_hitter_size=64, recent_size=444), 1, 1, 2, 122880, 128, torch.bfloat16),
    ("h2o_70k_f32", "h2o_l2", dict(start_size=4, heavy_hitter_size=64, recent_size=444), 1, 1, 2, 70000, 32, torch.float32),
    ("snapkv_120k_bf16", "snapkv_lite", dict(observation_window=32, keep_size=512), 1, 1, 1, 122880, 80, torch.bfloat16),
]


@pytest.mark.parametrize("name,method,kwargs,L,B,H,S,D,dtype", BIG, ids=[b[0] for b in BIG])
def test_selections_beyond_shared_memory_use_the_workspace(name, method, kwargs, L, B, H, S, D, dtype):
    kv = spread_cache(L, B, H, S, D, dtype)
    plans = plan_for(method, [S] * L, kwargs)
    lib = _engine.load_library()
    shape = _engine._SHAPE.pack(B, H, D, _engine.KVC_DTYPE[dtype], 0)
    packed = b"".join(_engine.PlanSet(plans).packed[i] for i in range(L))
    assert lib.kvc_workspace_bytes(shape, L, packed) > 0, "this case is meant to exceed the on-chip buffers"
    out, idx = _engine.run_plans(kv, plans, return_indices=True)
    api = kvcompress.get_compress_fn(method)(kv, **kwargs)
    for li, p in enumerate(plans):
        rows = idx[li]
        k_in, v_in = kv[li]
        assert torch.equal(out[li][0], gather_rows(k_in, rows)) and torch.equal(out[li][1], gather_rows(v_in, rows))
        assert torch.equal(api[li][0], out[li][0])
        sel = rows[..., p.sink:p.sink + p.k_sel] - p.sel_lo
        assert sel.min() >= 0 and sel.max() < p.sel_hi - p.sel_lo
        if method != "snapkv_lite":
            assert_valid_lowest(k_in[:, :, p.sel_lo:p.sel_hi], sel)
    # the in-place path takes the same fallback and keeps the same rows
    slab = kvcompress.KVSlabCache.from_legacy_cache(kv, capacity=S)
    slab.compress_(method, **kwargs)
    for li in range(L):
        assert torch.equal(slab[li][0], out[li][0]) and torch.equal(slab[li][1], out[li][1])


# ----------------------------------------------------------------------------------------------
# BASELINE.json FULL per-GPU sizes (not reduced batch): size-independent properties only
def _checksum(x: torch.Tensor) -> torch.Tensor:
    """Order-independent 64-bit checksum of the raw bytes of every row: [B,H,rows]."""
    raw = x.contiguous().view(torch.int16 if x.element_size() == 2 else torch.int32).to(torch.int64)
    weights = torch.arange(1, raw.size(-1) + 1, device=x.device, dtype=torch.int64) * 0x9E3779B1
    return (raw * weights).sum(-1)


FULL_SIZE = [
    # c2 at its full single-GPU size: 32 layers x (32,32,4096,80) bf16 = 42.9 GB
    ("c2_full", [("streaming_llm", dict(start_size=4, recent_size=508)),
                 ("fix_size_l2", dict(fix_kv_size=512, keep_ratio=0.2, strategy="keep_low"))], 32, 32, 32, 4096, 80),
    # c5's per-GPU shard on 8 GPUs: 32 layers x (8,8,32768,128) bf16 = 34.4 GB
    ("c5_shard", [("pyramid_kv", dict(base_size=512)), ("adaptive_l2", dict(target_size=512))], 32, 8, 8, 32768, 128),
    # c3's per-GPU shard on 8 GPUs (and one slab of the strong-scaled job): 32 layers x (32,32,8192,80) bf16 = 85.9 GB
    ("c3_shard", [("h2o_l2", dict(start_size=4, heavy_hitter_size=64, recent_size=444))], 32, 32, 32, 8192, 80),
    # c4 at its full single-GPU size: 32 layers x (16,8,32768,128) bf16 = 68.7 GB
    ("c4_full", [("snapkv_lite", dict(observation_window=32, keep_size=512, pooling_kernel=5))], 32, 16, 8, 32768, 128),
]


@pytest.mark.parametrize("name,calls,L,B,H,S,D", FULL_SIZE, ids=[f[0] for f in FULL_SIZE])
def test_full_baseline_size_checksums(name, calls, L, B, H, S, D):
    """Whole-job properties at the sizes bench.py runs: every output row is an input row (checksum of rows
    against the gather by the reported indices), sinks / tails are where the plan says, indices ascend,
    the kept rows are the lowest-norm rows, and a second call on the result is the identity."""
    import gc

    gc.collect()
    torch.cuda.empty_cache()   # blocks cached by earlier tests are not "free" to mem_get_info
    free, _ = torch.cuda.mem_get_info()
    need = 2 * L * B * H * S * D * 2 * 1.25
    if free < need:
        pytest.skip(f"needs {need / 2**30:.0f} GiB of free HBM")
    kv = spread_cache(L, B, H, S, D, torch.bfloat16)
    for method, kwargs in calls:
        plans = plan_for(method, [S] * L, kwargs)
        n0 = _engine.launch_count()
        out, idx = _engine.run_plans(kv, plans, return_indices=True)
        assert _engine.launch_count() - n0 == 1
        total_in = torch.zeros((), dtype=torch.int64, device="cuda")
        total_out = torch.zeros((), dtype=torch.int64, device="cuda")
        for li, p in enumerate(plans):
            if p.kind != P.GATHER:
                assert out[li][0] is kv[li][0]
                continue
            rows = idx[li].long()
            assert torch.all(rows[..., 1:] > rows[..., :-1])
            assert torch.equal(rows[..., :p.sink], torch.arange(p.sink, device="cuda").expand(B, H, -1))
            assert torch.equal(rows[..., p.sink + p.k_sel:], torch.arange(S - p.tail, S, device="cuda").expand(B, H, -1))
            for x_in, x_out in zip(kv[li], out[li]):
                want = torch.gather(_checksum(x_in), 2, rows)
                got = _checksum(x_out)
                assert torch.equal(got, want)
                total_in += want.sum()
                total_out += got.sum()
            if p.k_sel and li % 8 == 0 and p.score == P.SCORE_L2_LOW:
                # tie-aware selection rule on a sample of layers (fp64 norms are heavy)
                assert_valid_lowest(kv[li][0][:, :, p.sel_lo:p.sel_hi], idx[li][..., p.sink:p.sink + p.k_sel] - p.sel_lo)
            if p.k_sel and li % 8 == 0 and p.score == P.SCORE_SNAPKV_POOL:
                # snapkv_lite.py:96-134 with torch ops on this layer: the kept prefix rows are a top-k of the pooled
                # scores (ties at the threshold may differ: every kept score >= the k-th largest)
                norms = torch.norm(kv[li][0][:, :, :p.sel_hi], p=2, dim=-1)
                sc = (norms.max(dim=-1, keepdim=True)[0] + 1e-6) - norms
                pooled = torch.nn.functional.avg_pool1d(sc.reshape(B * H, 1, -1), p.pool_kernel, 1, p.pool_kernel // 2)
                pooled = pooled.reshape(B, H, -1)[..., :p.sel_hi]
                kth = torch.topk(pooled, p.k_sel, dim=-1)[0][..., -1:]
                picked = torch.gather(pooled, 2, idx[li][..., :p.k_sel].long())
                assert torch.all(picked >= kth)
        assert int(total_in) == int(total_out)  # checksum of checksums over the whole job
        again = kvcompress.get_compress_fn(method)(out, **kwargs)
        if method != "adaptive_l2":
            assert all(a[0] is b[0] for a, b in zip(again, out))
        del out, idx, again


# ----------------------------------------------------------------------------------------------
# §8f rows 2 and 3 against outputs of the real reference (tests/golden/make_extras_golden.py)
def test_evict_for_space_and_h2o_attention_manager_match_reference():
    import json
    import os

    import extras_cases as E

    want = json.load(open(os.path.join(os.path.dirname(cases.GOLDEN_NPZ), "extras_golden.json")))

    def rows_of(v):
        return v[0, :, :, 0].long().tolist()

    for name, seq_lens, num_coming, start, recent, skip in E.EVICT_CASES:
        kv = [(k.cuda(), v.cuda()) for k, v in E.evict_cache(seq_lens)]
        out = kvcompress.evict_for_space(kv, num_coming, start_size=start, recent_size=recent, skip_layers=skip)
        assert [k.size(2) for k, _ in out] == want["evict"][name]["lengths"], name
        assert [o[0] is i[0] for o, i in zip(out, kv)] == want["evict"][name]["untouched"], name
        assert [rows_of(v) for _, v in out] == want["evict"][name]["rows"], name
        for (k_in, _), (k_out, v_out) in zip(kv, out):
            assert torch.equal(k_out, gather_rows(k_in, v_out[..., 0].long()))

    c = E.H2O_CASE
    mgr = kvcompress.H2OAttentionManager(start_size=c["start_size"], heavy_hitter_size=c["heavy_hitter_size"],
                                         recent_size=c["recent_size"], num_layers=c["layers"], num_heads=c["heads"],
                                         decay_factor=c["decay_factor"])
    for step, ref in enumerate(want["h2o"]):
        kv, attn = E.h2o_inputs(step, ref["seq_len"])
        kv = [(k.cuda(), v.cuda()) for k, v in kv]
        out = kvcompress.h2o_attention_compress(kv, attention_scores=[a.cuda() for a in attn], h2o_manager=mgr,
                                                start_size=c["start_size"], heavy_hitter_size=c["heavy_hitter_size"],
                                                recent_size=c["recent_size"], skip_layers=c["skip_layers"])
        assert [k.size(2) for k, _ in out] == ref["lengths"], step
        assert [rows_of(v) for _, v in out] == ref["rows"], step


# ----------------------------------------------------------------------------------------------
# The compiled per-call binding (csrc/kvc_fast_binding.cpp) against the ctypes walk it replaces.
def test_compiled_binding_takes_the_common_call_and_matches_the_python_walk():
    assert _engine.fast_binding() is not None, "kvcompress/_kvc_fast*.so has not been built (g.build())"
    gen = torch.Generator(device="cuda").manual_seed(2)
    kv = [(torch.randn(2, 4, 900, 80, generator=gen, device="cuda").bfloat16(),
           torch.randn(2, 4, 900, 80, generator=gen, device="cuda").bfloat16()) for _ in range(5)]
    for method, kwargs in [("h2o_l2", dict(start_size=4, heavy_hitter_size=64, recent_size=444, skip_layers=[1])),
                           ("pyramid_kv", dict(base_size=512, layer_decay=0.7, min_size=64)),
                           ("streaming_llm", dict(start_size=4, recent_size=508)),
                           ("fix_size_l2", dict(fix_kv_size=512, keep_ratio=0.2, skip_layers=[0]))]:
        plans = _engine.PlanSet(plan_for(method, [900] * 5, kwargs))
        assert plans.fast() is not None
        n0 = _engine.launch_count()
        fast = plans.fast().run(kv, None)
        assert fast is not None and _engine.launch_count() - n0 == 1
        slow, _ = _engine.run_plans(kv, plans, return_indices=True)     # index output: the Python walk
        api = kvcompress.get_compress_fn(method)(kv, **kwargs)
        for li in range(5):
            if plans[li].kind == P.KEEP:
                assert fast[li] is kv[li] and api[li][0] is kv[li][0]
            else:
                assert torch.equal(fast[li][0], slow[li][0]) and torch.equal(fast[li][1], slow[li][1])
                assert torch.equal(api[li][0], slow[li][0]) and torch.equal(api[li][1], slow[li][1])
    # cases it leaves to the Python path: strided rows it cannot read in place, host tensors, wrong lengths
    odd = [(k[..., :72], v[..., :72]) for k, v in kv]                  # row stride 160 B, rows 144 B: fine in place
    plans = _engine.PlanSet(plan_for("streaming_llm", [900] * 5, dict(start_size=4, recent_size=508)))
    got = plans.fast().run(odd, None)
    assert got is not None and torch.equal(got[0][0], torch.cat([odd[0][0][:, :, :4], odd[0][0][:, :, -508:]], 2))
    misaligned = [(k[..., 4:76], v[..., 4:76]) for k, v in kv]         # rows start 8 bytes into a 16-byte chunk
    assert plans.fast().run(misaligned, None) is None
    want = torch.cat([misaligned[0][0][:, :, :4], misaligned[0][0][:, :, -508:]], 2)
    assert torch.equal(kvcompress.streaming_llm_compress(misaligned)[0][0], want)   # the Python path copies first
    assert plans.fast().run([(k[:, :, :800], v[:, :, :800]) for k, v in kv], None) is None


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_calls_on_another_gpu_leave_the_current_device_alone():
    """the library used to cudaSetDevice(shape->device) and leave it there."""
    assert torch.cuda.current_device() == 0
    kv = [(torch.randn(1, 2, 700, 80, device="cuda:1").bfloat16(), torch.randn(1, 2, 700, 80, device="cuda:1").bfloat16())]
    out = kvcompress.h2o_l2_compress(kv)
    idx_path, _ = _engine.run_plans(kv, plan_for("h2o_l2", [700], {}), return_indices=True)   # the ctypes walk
    torch.cuda.synchronize(1)
    assert torch.cuda.current_device() == 0
    assert out[0][0].device == torch.device("cuda", 1) and torch.equal(out[0][0], idx_path[0][0])
    assert torch.empty(1, device="cuda").device.index == 0
    slab = kvcompress.KVSlabCache.from_legacy_cache(kv, capacity=800)
    slab.compress_("h2o_l2")
    assert torch.cuda.current_device() == 0 and slab.device.index == 1


# ----------------------------------------------------------------------------------------------
# Random calls (the generator tests/test_oracle_live_reference.py runs against the live reference on CPU): random
# method, arguments, skip lists, ragged per-layer lengths, fp32 / bf16, continuous and tie-heavy keys.
@pytest.mark.parametrize("seed", range(32))
def test_random_calls_match_the_oracle(seed):
    from test_oracle_live_reference import draw_call

    rng = np.random.default_rng(9000 + seed)   # the same calls the live-reference test checks the oracle on
    for _ in range(4):
        method, kw, seq_lens, B, H, D, dtype, style = draw_call(rng)
        layers = cases.make_cache(int(rng.integers(0, 1 << 30)), seq_lens, B, H, D, dtype, style)
        what = (seed, method, kw, seq_lens, B, H, D, dtype, style)
        kv = kv_to_torch(layers, dtype)
        results = O.METHODS[method](layers, dtype, **kw)
        n0 = _engine.launch_count()
        out = kvcompress.get_compress_fn(method)(kv, **kw)
        plans = plan_for(method, seq_lens, kw)
        n_gather = sum(1 for p in plans if p.kind == P.GATHER)
        assert _engine.launch_count() - n0 == (1 if n_gather else 0), what
        assert [k.size(2) for k, _ in out] == O.out_lengths(layers, results), what
        out2, idx = _engine.run_plans(kv, plans, return_indices=True)
        for li, res in enumerate(results):
            k_in, v_in = kv[li]
            if res.untouched:
                assert out[li][0] is k_in and out[li][1] is v_in, (what, li)
                continue
            aliases = out[li][0].untyped_storage().data_ptr() == k_in.untyped_storage().data_ptr()
            assert aliases == bool(res.is_view), (what, li)   # tail-only results alias the input like the reference's
            assert res.is_view or (out[li][0].is_contiguous() and out[li][1].is_contiguous()), (what, li)
            if plans[li].kind != P.GATHER:
                rows = torch.from_numpy(res.rows).cuda()
                assert torch.equal(out[li][0], gather_rows(k_in, rows)) and torch.equal(out[li][1], gather_rows(v_in, rows))
                continue
            rows = idx[li]
            assert torch.equal(out2[li][0], gather_rows(k_in, rows)) and torch.equal(out2[li][1], gather_rows(v_in, rows))
            assert torch.equal(out2[li][0], out[li][0]) and torch.equal(out2[li][1], out[li][1]), (what, li)
            info = O.check_layer(layers[li][0], dtype, res, rows.cpu().numpy())
            assert info["valid"], (what, li, info)
            if dtype == "f32" and style != "ties":
                assert info["identical_heads"] == info["heads"], (what, li, info)


@pytest.mark.parametrize("method,kw", [
    ("fix_size_l2", dict(fix_kv_size=128, keep_ratio=0.2, skip_layers=[0])),
    ("h2o_l2", dict(start_size=4, heavy_hitter_size=32, recent_size=92)),
    ("streaming_llm", dict(start_size=4, recent_size=124)),
    ("snapkv_lite", dict(observation_window=16, keep_size=128)),
])
def test_function_call_replays_from_a_cuda_graph(method, kw):
    """The drop-in functions enqueue one launch on the current stream and synchronise nothing, so a steady-state call
    (same shapes every step, evaluate.py:154-166) can be captured with torch.cuda.graph and replayed: the replay reads
    the caches' CURRENT contents and writes the same output tensors."""
    L, B, H, S, D = 3, 2, 4, 129, 80
    gen = torch.Generator(device="cuda").manual_seed(3)
    kv = [(torch.randn(B, H, S, D, generator=gen, device="cuda").bfloat16(),
           torch.randn(B, H, S, D, generator=gen, device="cuda").bfloat16()) for _ in range(L)]
    fn = kvcompress.get_compress_fn(method)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):          # warm-up off the capture: library attributes, plan cache
        fn(kv, **kw)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    n0 = _engine.launch_count()
    with torch.cuda.graph(graph):
        out = fn(kv, **kw)
    assert _engine.launch_count() - n0 == 1
    for step in range(3):
        for k, v in kv:                    # new cache contents in the same tensors
            k.copy_(torch.randn(B, H, S, D, generator=gen, device="cuda") * (1 + step))
            v.copy_(torch.randn(B, H, S, D, generator=gen, device="cuda"))
        graph.replay()
        want = fn(kv, **kw)
        for li in range(L):
            if want[li][0] is kv[li][0]:
                assert out[li][0] is kv[li][0]
                continue
            assert torch.equal(out[li][0], want[li][0]) and torch.equal(out[li][1], want[li][1]), (method, step, li)


def test_concurrent_calls_from_two_threads_on_their_own_streams():
    """The library keeps no mutable state besides per-thread error text and a launch counter: two Python threads, each
    on its own CUDA stream, compress their own caches at the same time and get what a serial call gets."""
    import threading

    L, B, H, S, D = 4, 2, 4, 700, 80
    gen = torch.Generator(device="cuda").manual_seed(99)
    caches = [[(torch.randn(B, H, S, D, generator=gen, device="cuda").bfloat16(),
                torch.randn(B, H, S, D, generator=gen, device="cuda").bfloat16()) for _ in range(L)] for _ in range(2)]
    calls = [("h2o_l2", dict(start_size=4, heavy_hitter_size=32, recent_size=92)),
             ("snapkv_lite", dict(observation_window=16, keep_size=128))]
    want = [kvcompress.get_compress_fn(m)(kv, **kw) for kv, (m, kw) in zip(caches, calls)]
    torch.cuda.synchronize()
    errors, results = [], [None, None]

    def worker(i):
        try:
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                m, kw = calls[i]
                for _ in range(40):
                    out = kvcompress.get_compress_fn(m)(caches[i], **kw)
                stream.synchronize()
            results[i] = out
        except Exception as exc:  # surfaced below: an exception in a thread would otherwise be lost
            errors.append(repr(exc))

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for i in range(2):
        for li in range(L):
            assert torch.equal(results[i][li][0], want[i][li][0]) and torch.equal(results[i][li][1], want[i][li][1]), (i, li)
