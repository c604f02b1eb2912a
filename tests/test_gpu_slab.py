"""-m gpu: KVSlabCache (in-place append + in-place compression) against the
out-of-place functions walked through the same decode loop — the reference's loop shape
(evaluate.py:132-166): append one token per layer, compress, repeat."""

import pytest
import torch

import kvcompress
from kvcompress import KVSlabCache, _engine
from kvcompress import _planner as P
from test_planner import plan_for

pytestmark = pytest.mark.gpu

DT = {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32}


def rand_rows(B, H, T, D, dtype, gen, spread=True):
    k = torch.randn(B, H, T, D, generator=gen, device="cuda")
    if spread:
        k = k * torch.exp(0.35 * torch.randn(B, H, T, 1, generator=gen, device="cuda"))
    v = torch.randn(B, H, T, D, generator=gen, device="cuda")
    return k.to(dtype), v.to(dtype)


LOOPS = [
    ("streaming_llm", dict(start_size=4, recent_size=60), "bf16", 80),
    ("streaming_llm", dict(start_size=0, recent_size=33, skip_layers=[1]), "f32", 128),
    ("h2o_l2", dict(start_size=4, heavy_hitter_size=16, recent_size=44), "bf16", 80),
    ("h2o_l2", dict(start_size=4, heavy_hitter_size=16, recent_size=44), "f32", 80),
    ("fix_size_l2", dict(fix_kv_size=64, keep_ratio=0.2, skip_layers=[0]), "bf16", 128),
    ("fix_size_l2", dict(fix_kv_size=64, keep_ratio=0.5, strategy="keep_high", skip_layers=[]), "f16", 64),
    ("fix_size_l2", dict(fix_kv_size=64, keep_ratio=0.25, strategy="random", skip_layers=[]), "bf16", 80),
    ("fix_size_l2", dict(fix_kv_size=64, keep_ratio=1.0, skip_layers=[]), "bf16", 80),   # tail-only (a view in the reference)
    ("snapkv_lite", dict(observation_window=8, keep_size=64, pooling_kernel=5), "bf16", 128),
    ("snapkv_lite", dict(observation_window=8, keep_size=64, pooling_kernel=4), "f32", 80),
    ("pyramid_kv", dict(base_size=64, layer_decay=0.8, min_size=16), "bf16", 80),
    ("adaptive_l2", dict(target_size=64, soft_limit=32, hard_limit=100), "bf16", 128),
    ("l2_compress", dict(keep_ratio=0.9, prune_after=60, skip_layers=[0]), "f32", 80),
    ("recent_only", dict(window_size=50, skip_layers=[0]), "bf16", 80),
]


@pytest.mark.parametrize("method,kwargs,dtype,D", LOOPS, ids=[f"{m}-{i}" for i, (m, *_r) in enumerate(LOOPS)])
def test_decode_loop_matches_out_of_place(method, kwargs, dtype, D):
    """Same lengths and bit-identical K/V after every step of prefill -> (append 1 token, compress) x N."""
    L, B, H, S0, steps = 3, 2, 3, 150, 24
    dt = DT[dtype]
    gen = torch.Generator(device="cuda").manual_seed(11)
    fn = kvcompress.get_compress_fn(method)
    prefill = [rand_rows(B, H, S0, D, dt, gen) for _ in range(L)]
    slab = KVSlabCache.from_legacy_cache(prefill, capacity=256)
    kv = [(k.clone(), v.clone()) for k, v in prefill]
    for (k, v), (ks, vs) in zip(kv, slab):
        assert torch.equal(k, ks) and torch.equal(v, vs)
    for step in range(steps):
        torch.manual_seed(1000 + step)
        kv = fn(kv, **kwargs)
        torch.manual_seed(1000 + step)
        n0 = _engine.launch_count()
        slab.compress_(method, **kwargs)
        assert _engine.launch_count() - n0 <= 1, "in-place compression of all layers is ONE launch"
        assert [k.size(2) for k, _ in kv] == slab.lengths, (step, slab.lengths)
        for li, ((k, v), (ks, vs)) in enumerate(zip(kv, slab)):
            assert torch.equal(k, ks), (method, step, li)
            assert torch.equal(v, vs), (method, step, li)
        new = [rand_rows(B, H, 1, D, dt, gen) for _ in range(L)]
        kv = [(torch.cat([k, nk], 2), torch.cat([v, nv], 2)) for (k, v), (nk, nv) in zip(kv, new)]
        if step % 3 == 1:
            slab.append(new)                       # all layers, one launch
        elif step % 3 == 2 and len(set(slab.lengths)) == 1:
            slab.append_stacked(torch.stack([k for k, _ in new]), torch.stack([v for _, v in new]))
        else:
            for li, (nk, nv) in enumerate(new):    # HF-style per-layer update
                full_k, full_v = slab.update(nk, nv, li)
                assert full_k.size(2) == kv[li][0].size(2)


@pytest.mark.parametrize("dtype", ["bf16", "f16", "f32"])
@pytest.mark.parametrize("D", [64, 80, 96, 128])
def test_append_copies_rows_and_records_torch_norms(dtype, D):
    dt = DT[dtype]
    if D * torch.empty((), dtype=dt).element_size() // 16 not in (8, 10, 12, 16, 20, 32):
        pytest.skip("row width not covered by the slab kernels")
    gen = torch.Generator(device="cuda").manual_seed(3)
    L, B, H = 2, 2, 4
    slab = KVSlabCache(L, B, H, D, 300, dt)
    ref = [[], []]
    for T in (130, 1, 7, 1):
        new = [rand_rows(B, H, T, D, dt, gen) for _ in range(L)]
        # strided new rows ([B,T,H,D] storage, as attention layers produce them)
        new = [(k.permute(0, 2, 1, 3).contiguous().permute(0, 2, 1, 3), v) for k, v in new]
        slab.append(new)
        for li in range(L):
            ref[li].append(new[li])
    for li in range(L):
        k = torch.cat([x[0] for x in ref[li]], 2)
        v = torch.cat([x[1] for x in ref[li]], 2)
        assert slab.lengths[li] == 139
        assert torch.equal(slab[li][0], k) and torch.equal(slab[li][1], v)
        want = torch.linalg.vector_norm(k.float(), dim=-1)
        got = slab.key_norms(li).float()
        tol = 1e-6 if dtype == "f32" else (2 ** -8 if dtype == "bf16" else 2 ** -11)
        assert torch.all((got - want).abs() <= tol * want + 1e-30)
        if dtype != "f32":  # == torch.norm's rounding except where the fp32 sum order flips the last rounding
            assert (got == want.to(dt).float()).float().mean() > 0.995


def test_in_place_indices_match_out_of_place_and_functions_accept_slabs():
    gen = torch.Generator(device="cuda").manual_seed(5)
    L, B, H, S, D = 4, 2, 8, 2000, 128
    kv = [rand_rows(B, H, S, D, torch.bfloat16, gen) for _ in range(L)]
    for method, kwargs in (("h2o_l2", dict(start_size=4, heavy_hitter_size=64, recent_size=444)),
                           ("snapkv_lite", dict(observation_window=32, keep_size=512)),
                           ("pyramid_kv", dict(base_size=512)),
                           ("adaptive_l2", dict(target_size=512))):
        slab = KVSlabCache.from_legacy_cache(kv, capacity=2048)
        # a registered function applied to the slab (through to_legacy_cache views)
        out = kvcompress.get_compress_fn(method)(slab, **kwargs)
        plans = plan_for(method, [S] * L, kwargs)
        out2, idx = _engine.run_plans(kv, plans, return_indices=True)
        _, idx_ip = slab.compress_(method, return_indices=True, **kwargs)
        for li in range(L):
            assert torch.equal(out[li][0], out2[li][0]) and torch.equal(out[li][1], out2[li][1])
            assert torch.equal(slab[li][0], out2[li][0]) and torch.equal(slab[li][1], out2[li][1])
            if plans[li].kind == P.GATHER:
                assert torch.equal(idx_ip[li], idx[li])
                # norms slid with their rows
                want = torch.gather(torch.linalg.vector_norm(kv[li][0].float(), dim=-1).to(torch.bfloat16), 2,
                                    idx[li].long())
                assert (slab.key_norms(li) == want).float().mean() > 0.995


def test_slab_errors():
    slab = KVSlabCache(1, 1, 2, 80, 32, torch.bfloat16)
    k = torch.zeros(1, 2, 40, 80, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(ValueError, match="exceed the slab capacity"):
        slab.update(k, k, 0)
    with pytest.raises(ValueError, match="must be torch.bfloat16"):
        slab.update(k[:, :, :4].float(), k[:, :, :4].float(), 0)
    with pytest.raises(RuntimeError, match="must live on"):
        slab.update(k[:, :, :4].cpu(), k[:, :, :4].cpu(), 0)
    with pytest.raises(ValueError, match="Unknown method"):
        slab.compress_("nope")
    with pytest.raises(ValueError, match="multiple of 16 bytes"):
        KVSlabCache(1, 1, 2, 12, 32, torch.bfloat16)       # 24-byte rows
    with pytest.raises(RuntimeError, match="no CPU path"):
        KVSlabCache(1, 1, 2, 80, 32, torch.bfloat16, device="cpu")


def test_unit_pitch_avoids_channel_aligned_strides():
    """A (batch, head) pitch that is a multiple of 16 KB lines the same rows of every unit up on the same HBM channels
    (profiles/r02_stream_copy_control.json): the slab pads such capacities, the valid rows stay where they were."""
    slab = KVSlabCache(1, 1, 2, 80, 4096, torch.bfloat16)          # 4096 x 160 B = 640 KB
    assert slab.pitch == 4104 and slab.k.shape == (1, 1, 2, 4096, 80) and slab.k.stride(2) == 4104 * 80
    assert KVSlabCache(1, 1, 2, 128, 32768, torch.bfloat16).pitch == 32776
    assert KVSlabCache(1, 1, 2, 80, 4104, torch.bfloat16).pitch == 4104 and KVSlabCache(1, 1, 2, 80, 513, torch.bfloat16).pitch == 513


def test_captured_decode_step_replays_append_and_compress():
    """CUDA-graph replay of (append one token, compress in place) == the same two calls made eagerly."""
    gen = torch.Generator(device="cuda").manual_seed(9)
    L, B, H, D, cap = 4, 2, 4, 80, 64
    kw = dict(start_size=4, heavy_hitter_size=16, recent_size=44)
    prefill = [rand_rows(B, H, cap, D, torch.bfloat16, gen) for _ in range(L)]
    a = KVSlabCache.from_legacy_cache(prefill, capacity=cap + 8)
    b = KVSlabCache.from_legacy_cache(prefill, capacity=cap + 8)
    step = a.capture_step("h2o_l2", skip_layers=[], **kw)
    n0 = _engine.launch_count()
    for _ in range(12):
        k_new = torch.randn(L, B, H, 1, D, generator=gen, device="cuda").bfloat16()
        v_new = torch.randn(L, B, H, 1, D, generator=gen, device="cuda").bfloat16()
        step.k_new.copy_(k_new)
        step.v_new.copy_(v_new)
        step()
        b.append_stacked(k_new, v_new).compress_("h2o_l2", skip_layers=[], **kw)
        assert a.lengths == b.lengths == [cap] * L
        for li in range(L):
            assert torch.equal(a[li][0], b[li][0]) and torch.equal(a[li][1], b[li][1])
    assert _engine.launch_count() - n0 == 24  # only the eager slab launched through the library during the loop
    with pytest.raises(ValueError, match="steady"):
        KVSlabCache.from_legacy_cache([(k[:, :, :40], v[:, :, :40]) for k, v in prefill], capacity=cap + 8) \
            .capture_step("h2o_l2", skip_layers=[], **kw)


# ----------------------------------------------------------------------------------------------
# Stored norms: functions called on a slab cache rank rows from the norms recorded at append time, and a slab in
# pinned HOST memory (offloaded cache) gives the same bytes as the device slab and as the plain (K, V) functions.
STORED = [
    ("fix_size_l2", dict(fix_kv_size=128, keep_ratio=0.2, skip_layers=[0]), "bf16", 80),
    ("fix_size_l2", dict(fix_kv_size=128, keep_ratio=0.3, strategy="keep_high", skip_layers=[]), "f32", 128),
    ("h2o_l2", dict(start_size=4, heavy_hitter_size=32, recent_size=92), "bf16", 128),
    ("snapkv_lite", dict(observation_window=16, keep_size=128, pooling_kernel=5), "bf16", 80),
    ("snapkv_lite", dict(observation_window=16, keep_size=128, pooling_kernel=4), "f16", 72),
    ("pyramid_kv", dict(base_size=128, layer_decay=0.8, min_size=32), "bf16", 128),
    ("adaptive_l2", dict(target_size=128, soft_limit=64, hard_limit=300), "bf16", 80),
    ("l2_compress", dict(keep_ratio=0.7, prune_after=100, skip_layers=[1]), "f32", 80),
    ("streaming_llm", dict(start_size=4, recent_size=124), "bf16", 80),
    ("fix_size_l2", dict(fix_kv_size=128, keep_ratio=0.25, strategy="random", skip_layers=[]), "bf16", 80),
]


@pytest.mark.parametrize("method,kwargs,dtype,D", STORED, ids=[f"{m}-{i}" for i, (m, *_r) in enumerate(STORED)])
def test_pinned_slab_equals_device_slab_equals_functions(method, kwargs, dtype, D):
    L, B, H, S = 3, 2, 3, 700
    dt = DT[dtype]
    gen = torch.Generator(device="cuda").manual_seed(5)
    kv = [rand_rows(B, H, S, D, dt, gen) for _ in range(L)]
    fn = kvcompress.get_compress_fn(method)

    def run(x):
        torch.manual_seed(77)  # strategy="random" draws from the default generators
        return fn(x, **kwargs)

    want = run(kv)                                              # plain (K, V) list: the K scan
    dev = KVSlabCache.from_legacy_cache(kv, capacity=S + 3)
    host = KVSlabCache.from_legacy_cache(kv, capacity=S + 3, pinned=True)
    assert host.pinned and host.k.is_pinned() and not host.k.is_cuda and host.n.is_pinned()
    for l in range(L):
        assert torch.equal(host.key_norms(l).cuda(), dev.key_norms(l))
        assert torch.equal(host[l][0].cuda(), kv[l][0]) and torch.equal(host[l][1].cuda(), kv[l][1])
    random = kwargs.get("strategy") == "random"
    # out of place, scores from the stored norms (device slab, then the host-resident one: outputs land in pinned memory)
    n0 = _engine.launch_count()
    got_dev = run(dev)
    assert _engine.launch_count() - n0 <= 1
    got_host = run(host)
    for li in range(L):
        assert torch.equal(got_dev[li][0], want[li][0]) and torch.equal(got_dev[li][1], want[li][1]), (method, li)
        if want[li][0] is not kv[li][0]:
            assert not got_host[li][0].is_cuda and got_host[li][0].is_pinned()
        if not random:   # the host slab draws its random rows from the CPU generator, like the reference on CPU tensors
            assert torch.equal(got_host[li][0].cuda(), want[li][0]) and torch.equal(got_host[li][1].cuda(), want[li][1])
        assert got_host[li][0].shape == want[li][0].shape
    if not random:
        # the torch idioms for host destinations: queue now, synchronise later; or land the result on the GPU
        queued = fn(host, non_blocking=True, **kwargs)
        fetched = fn(host, output_device="cuda", **kwargs)
        torch.cuda.synchronize()
        for li in range(L):
            assert torch.equal(queued[li][0].cuda(), want[li][0]) and torch.equal(queued[li][1].cuda(), want[li][1])
            if want[li][0] is not kv[li][0]:
                assert fetched[li][0].is_cuda and torch.equal(fetched[li][0], want[li][0]) and torch.equal(fetched[li][1], want[li][1])
    # in place
    torch.manual_seed(77)
    dev.compress_(method, **kwargs)
    torch.manual_seed(77)
    host.compress_(method, **kwargs)
    assert dev.lengths == host.lengths == [k.size(2) for k, _ in want]
    for li in range(L):
        assert torch.equal(dev[li][0], want[li][0]) and torch.equal(dev[li][1], want[li][1])
        if not random:
            assert torch.equal(host[li][0].cuda(), want[li][0]) and torch.equal(host[li][1].cuda(), want[li][1])
            assert torch.equal(host.key_norms(li).cuda(), dev.key_norms(li))
    # the host slab keeps decoding: append a token from the device, compress again
    new = [rand_rows(B, H, 1, D, dt, gen) for _ in range(L)]
    dev.append(new)
    host.append(new)
    if not random:
        for li in range(L):
            assert torch.equal(host[li][0].cuda(), dev[li][0]) and torch.equal(host.key_norms(li).cuda(), dev.key_norms(li))


@pytest.mark.parametrize("S,keep_ratio", [(4000, 0.8), (2600, 0.79), (9000, 0.3)])
def test_in_place_with_more_kept_rows_than_histogram_bins(S, keep_ratio):
    """The in-place plan lets the kept-index list alias the (dead) 2048-bin histogram; selections that keep more rows
    than that get their own list.  Both sides of the threshold (k_sel = 3200 / 2054 / 2700) against the function."""
    L, B, H, D = 2, 2, 3, 80
    gen = torch.Generator(device="cuda").manual_seed(S)
    kv = [rand_rows(B, H, S, D, torch.float32, gen) for _ in range(L)]
    kw = dict(keep_ratio=keep_ratio, prune_after=100, skip_layers=[])
    want = kvcompress.l2_compress(kv, **kw)
    slab = KVSlabCache.from_legacy_cache(kv, capacity=S + 4)
    slab, idx = slab.compress_("l2_compress", return_indices=True, **kw)
    fresh = KVSlabCache.from_legacy_cache(want, capacity=S + 4)   # the norms the append kernel records for the kept rows
    for l in range(L):
        assert slab.lengths[l] == want[l][0].size(2)
        assert torch.equal(slab[l][0], want[l][0]) and torch.equal(slab[l][1], want[l][1])
        rows = idx[l].long()
        assert torch.all(rows[..., 1:] > rows[..., :-1])
        assert torch.equal(slab.key_norms(l), fresh.key_norms(l))  # norms slid with their rows


def test_h2o_attention_manager_in_place_matches_the_reference_rows():
    """``compress_("h2o_attention", h2o_manager=...)``: the manager's heavy hitters as caller-supplied rows of the
    in-place compaction, against the rows the REAL reference kept (tests/golden/extras_golden.json) and against
    the function; without a manager it is h2o_l2."""
    import json
    import os

    import cases
    import extras_cases as E

    want = json.load(open(os.path.join(os.path.dirname(cases.GOLDEN_NPZ), "extras_golden.json")))
    c = E.H2O_CASE
    kw = dict(start_size=c["start_size"], heavy_hitter_size=c["heavy_hitter_size"], recent_size=c["recent_size"],
              skip_layers=c["skip_layers"])

    def manager():
        return kvcompress.H2OAttentionManager(start_size=c["start_size"], heavy_hitter_size=c["heavy_hitter_size"],
                                              recent_size=c["recent_size"], num_layers=c["layers"], num_heads=c["heads"],
                                              decay_factor=c["decay_factor"])

    m_fn, m_slab = manager(), manager()
    for step, ref in enumerate(want["h2o"]):
        kv, attn = E.h2o_inputs(step, ref["seq_len"])
        kv = [(k.cuda(), v.cuda()) for k, v in kv]
        attn = [a.cuda() for a in attn]
        out = kvcompress.h2o_attention_compress(kv, attention_scores=attn, h2o_manager=m_fn, **kw)
        slab = KVSlabCache.from_legacy_cache(kv, capacity=ref["seq_len"] + 4)
        n0 = _engine.launch_count()
        slab.compress_("h2o_attention", attention_scores=attn, h2o_manager=m_slab, **kw)
        assert _engine.launch_count() - n0 <= 1
        assert slab.lengths == ref["lengths"], step
        for l in range(len(kv)):
            assert slab[l][1][0, :, :, 0].long().tolist() == ref["rows"][l], (step, l)   # V carries the row positions
            assert torch.equal(slab[l][0], out[l][0]) and torch.equal(slab[l][1], out[l][1]), (step, l)
    # no manager: the h2o_l2 selection (reference h2o_attention.py:337-351)
    gen = torch.Generator(device="cuda").manual_seed(8)
    kv = [rand_rows(2, 3, 300, 80, torch.bfloat16, gen) for _ in range(2)]
    a = KVSlabCache.from_legacy_cache(kv, capacity=304).compress_("h2o_attention", start_size=4, heavy_hitter_size=16, recent_size=44)
    b = KVSlabCache.from_legacy_cache(kv, capacity=304).compress_("h2o_l2", start_size=4, heavy_hitter_size=16, recent_size=44)
    for l in range(2):
        assert torch.equal(a[l][0], b[l][0]) and torch.equal(a[l][1], b[l][1])


def test_chunked_prefill_evicts_in_place_like_the_function():
    """evict_for_space before every prefill chunk (reference streaming_llm.py:114-170), on the slab in place and with
    the function on plain (K, V) lists: the same cache after every chunk."""
    L, B, H, D, chunk = 3, 2, 3, 80, 48
    gen = torch.Generator(device="cuda").manual_seed(21)
    slab = KVSlabCache(L, B, H, D, capacity=160, dtype=torch.bfloat16)
    kv = None
    for step in range(7):
        new = [rand_rows(B, H, chunk, D, torch.bfloat16, gen) for _ in range(L)]
        if kv is not None:
            kv = kvcompress.evict_for_space(kv, chunk, start_size=4, recent_size=100, skip_layers=[1] if step < 3 else [])
            n0 = _engine.launch_count()
            slab.evict_for_space_(chunk, start_size=4, recent_size=100, skip_layers=[1] if step < 3 else [])
            assert _engine.launch_count() - n0 <= 1
            if step >= 3:   # layer 1 was skipped for three chunks (it grew); from now on it is evicted like the others
                assert slab.lengths[1] + chunk <= 160
        kv = new if kv is None else [(torch.cat([k, nk], 2), torch.cat([v, nv], 2)) for (k, v), (nk, nv) in zip(kv, new)]
        slab.append(new)
        assert slab.lengths == [k.size(2) for k, _ in kv], step
        for l in range(L):
            assert torch.equal(slab[l][0], kv[l][0]) and torch.equal(slab[l][1], kv[l][1]), (step, l)
            assert torch.equal(slab.key_norms(l), torch.norm(kv[l][0], p=2, dim=-1)), (step, l)


def test_host_rows_the_kernels_cannot_read_in_place_are_re_pinned():
    """Pinned host rows whose layout the kernels cannot read in place (last dimension not dense) are re-laid-out into a
    PINNED temporary (``.contiguous()`` alone gives pageable memory the GPU cannot reach) and the launch finishes before
    that temporary is released — on a device slab's append paths and on the functions' ``non_blocking`` / ``output_device`` forms."""
    L, B, H, S, D = 2, 2, 3, 300, 80
    gen = torch.Generator(device="cuda").manual_seed(11)
    kv = [rand_rows(B, H, S, D, torch.bfloat16, gen) for _ in range(L)]
    # [B,H,D,S] storage viewed as [B,H,S,D]: stride(3) != 1
    odd = [tuple(t.cpu().transpose(2, 3).contiguous().pin_memory().transpose(2, 3) for t in pair) for pair in kv]
    assert odd[0][0].is_pinned() and odd[0][0].stride(3) != 1
    dev = KVSlabCache(L, B, H, D, capacity=S + 8, dtype=torch.bfloat16)
    dev.append(odd)                                   # every layer in one launch, from host temporaries
    for l in range(L):
        assert torch.equal(dev[l][0], kv[l][0]) and torch.equal(dev[l][1], kv[l][1])
        assert torch.equal(dev.key_norms(l), torch.norm(kv[l][0], p=2, dim=-1))
    one = [tuple(t.cpu().transpose(2, 3).contiguous().pin_memory().transpose(2, 3) for t in rand_rows(B, H, 1, D, torch.bfloat16, gen))
           for _ in range(L)]
    for l in range(L):
        k_all, _ = dev.update(one[l][0], one[l][1], l)  # the per-layer HF path
        assert torch.equal(k_all[:, :, -1:], one[l][0].cuda())
    kw = dict(fix_kv_size=64, keep_ratio=0.25, skip_layers=[])
    want = kvcompress.fix_size_l2_compress(kv, **kw)
    queued = kvcompress.fix_size_l2_compress(odd, non_blocking=True, **kw)
    fetched = kvcompress.fix_size_l2_compress(odd, output_device="cuda", **kw)
    torch.cuda.synchronize()
    for l in range(L):
        assert torch.equal(queued[l][0].cuda(), want[l][0]) and torch.equal(queued[l][1].cuda(), want[l][1])
        assert torch.equal(fetched[l][0], want[l][0]) and torch.equal(fetched[l][1], want[l][1])


def test_stored_norms_replace_the_scan_bytes():
    """With stored norms no K row of the selection region is read for scoring: poison every K row that is NOT kept
    after recording the norms — the function must still return the rows the scan would have kept."""
    L, B, H, S, D = 2, 2, 2, 900, 80
    gen = torch.Generator(device="cuda").manual_seed(9)
    kv = [rand_rows(B, H, S, D, torch.bfloat16, gen) for _ in range(L)]
    want, idx = _engine.run_plans(kv, plan_for("h2o_l2", [S] * L, dict(start_size=4, heavy_hitter_size=32, recent_size=92)),
                                  return_indices=True)
    slab = KVSlabCache.from_legacy_cache(kv, capacity=S)
    for li in range(L):
        kept = torch.zeros(B, H, S, dtype=torch.bool, device="cuda").scatter_(2, idx[li].long(), True)
        slab.k[li][~kept] = float("nan")                       # the norms were recorded at append time
    got = kvcompress.h2o_l2_compress(slab, start_size=4, heavy_hitter_size=32, recent_size=92)
    for li in range(L):
        assert torch.equal(got[li][0], want[li][0]) and torch.equal(got[li][1], want[li][1])


# ----------------------------------------------------------------------------------------------
# Soak: randomised plans through the three data paths (K scan, stored norms out of place, in place), every result
# compared bit for bit with torch ops on the same norms, every launch repeated.  compute-sanitizer is closed on this GPU
# pool (profiles/r02_sanitizer.md); this is the standing check on the mbarrier / bulk-copy / in-place-overlap code:
# a race shows up as a mismatch or as two launches that disagree.
def _expected_rows(norms, plan):
    """Kept rows per (b, h): sinks, the k_sel lowest (highest) norms of [sel_lo, sel_hi) with ties to the lowest index,
    the tail — torch ops only."""
    B, H, S = norms.shape
    region = norms[:, :, plan.sel_lo:plan.sel_hi].float()
    if plan.score == P.SCORE_L2_HIGH:
        region = -region
    order = torch.sort(region, dim=-1, stable=True)[1][..., :plan.k_sel]
    sel = torch.sort(order, dim=-1)[0] + plan.sel_lo
    dev = norms.device
    sink = torch.arange(plan.sink, device=dev).expand(B, H, -1)
    tail = torch.arange(S - plan.tail, S, device=dev).expand(B, H, -1)
    return torch.cat([sink, sel, tail], dim=-1)


@pytest.mark.parametrize("seed", range(6))
def test_soak_random_plans_three_paths_agree_with_torch(seed):
    import random

    rng = random.Random(1000 + seed)
    gen = torch.Generator(device="cuda").manual_seed(seed)
    for _case in range(20):
        dtype = rng.choice(["bf16", "bf16", "f16", "f32"])
        D = rng.choice([64, 80, 128, 72, 40, 96] if dtype != "f32" else [32, 80, 100, 128])
        L, B, H = rng.randint(1, 4), rng.randint(1, 3), rng.randint(1, 4)
        plans = []
        lens = []
        for _ in range(L):
            S = rng.choice([rng.randint(40, 300), rng.randint(300, 2500), rng.randint(2500, 6000)])
            sink = rng.choice([0, 0, 4, rng.randint(0, min(40, S // 4))])
            tail = rng.choice([0, rng.randint(0, S // 3), rng.randint(0, min(600, S // 2))])
            lo = sink + rng.choice([0, 0, rng.randint(0, 10)])
            hi = S - tail - rng.choice([0, 0, rng.randint(0, 10)])
            if hi < lo:
                lo = hi = sink
            k = rng.choice([0, 1, rng.randint(0, hi - lo), (hi - lo) // 2, hi - lo])
            score = rng.choice([P.SCORE_L2_LOW, P.SCORE_L2_LOW, P.SCORE_L2_HIGH]) if k else P.SCORE_NONE
            if sink + k + tail == 0:
                tail = 1
                hi = min(hi, S - 1)
                lo = min(lo, hi)
                k = min(k, hi - lo)
            plans.append(P.LayerPlan(P.GATHER, S, sink, lo, hi, k, tail, score, 1))
            lens.append(S)
        dt = DT[dtype]
        kv = []
        for S in lens:
            k, v = rand_rows(B, H, S, D, dt, gen, spread=rng.random() < 0.7)
            if rng.random() < 0.3:                       # ties: rows drawn from a small codebook
                book = k[:, :, :7].clone()
                k = book[:, :, torch.randint(0, 7, (S,), generator=gen, device="cuda")]
            kv.append((k.contiguous(), v))
        slab = KVSlabCache.from_legacy_cache(kv, capacity=max(lens) + rng.choice([0, 3, 8]))
        want = []
        for li, p in enumerate(plans):
            rows = _expected_rows(slab.key_norms(li), p)
            ix = rows.unsqueeze(-1).expand(-1, -1, -1, D)
            want.append((torch.gather(kv[li][0], 2, ix), torch.gather(kv[li][1], 2, ix), rows))
        ps = _engine.PlanSet(plans)
        scan1, idx1 = _engine.run_plans(kv, ps, return_indices=True)                    # K scan (ctypes walk)
        scan2 = _engine.run_plans(kv, ps)                                               # K scan (compiled binding)
        by_norms = _engine.run_plans(slab.to_legacy_cache(), ps, norms=slab.key_norm_layers())
        by_norms2 = _engine.run_plans(slab.to_legacy_cache(), ps, norms=slab.key_norm_layers())
        slab.apply_plans_(ps)                                                           # in place
        for li, (wk, wv, rows) in enumerate(want):
            tag = (seed, _case, li, dtype, D, plans[li])
            assert torch.equal(idx1[li].long(), rows), tag
            for got in (scan1[li], scan2[li], by_norms[li], by_norms2[li], slab[li]):
                assert torch.equal(got[0], wk) and torch.equal(got[1], wv), tag
            assert slab.lengths[li] == rows.size(-1)
            assert torch.equal(slab.key_norms(li), torch.gather(KVSlabCache.from_legacy_cache([kv[li]], capacity=lens[li]).key_norms(0), 2, rows)), tag
