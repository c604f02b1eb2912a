"""-m gpu: the opt-in tcgen05 q.K^T observation-window vote (an extension: the reference's snapkv_lite has no
queries) against a plain torch fp32 reference of the same op, and its select against the reference scores."""

import math

import pytest
import torch

import kvcompress
from kvcompress import _engine

pytestmark = pytest.mark.gpu


def vote_reference(keys, queries, window):
    """fp32: softmax over all S keys (causal inside the window), summed over window queries and group heads."""
    B, H, S, D = keys.shape
    G = queries.size(1) // H
    P = S - window
    k = keys.float().repeat_interleave(G, dim=1)                       # [B, H*G, S, D]
    s = torch.matmul(queries.float(), k.transpose(-1, -2)) / math.sqrt(D)   # [B, H*G, W, S]
    pos_q = P + torch.arange(window, device=keys.device).view(1, 1, window, 1)
    pos_k = torch.arange(S, device=keys.device).view(1, 1, 1, S)
    s = s.masked_fill(pos_k > pos_q, float("-inf"))
    attn = torch.softmax(s, dim=-1)
    return attn[..., :P].sum(dim=2).view(B, H, G, P).sum(dim=2)        # [B, H, P]


CASES = [
    # B, H, G, W, S, D, dtype
    (1, 2, 4, 32, 1000, 128, torch.bfloat16),   # Llama-3-8B GQA shape: 4 query heads x 32 window rows = 128
    (2, 3, 1, 32, 700, 80, torch.bfloat16),     # Pythia MHA: 32 query rows, 96 padding rows
    (1, 2, 2, 16, 300, 64, torch.float16),
    (1, 1, 4, 32, 129, 128, torch.bfloat16),    # P = 97: a single partial tile, window straddles it
    (2, 8, 4, 32, 4096, 128, torch.bfloat16),
    (1, 4, 2, 32, 1500, 128, torch.bfloat16),   # 64 query rows: two replicas of a 64-row block
    (2, 2, 3, 16, 900, 80, torch.float16),      # 48 query rows: 64-row blocks with 16 padding rows each
    (1, 3, 1, 8, 400, 64, torch.bfloat16),      # 8 query rows: four replicas of a 32-row block, 24 padding rows
]


@pytest.mark.parametrize("B,H,G,W,S,D,dtype", CASES, ids=[f"B{c[0]}H{c[1]}G{c[2]}W{c[3]}S{c[4]}D{c[5]}" for c in CASES])
def test_votes_match_torch_reference(B, H, G, W, S, D, dtype):
    gen = torch.Generator(device="cuda").manual_seed(S + D)
    keys = [torch.randn(B, H, S, D, generator=gen, device="cuda").to(dtype) for _ in range(2)]
    qs = [(1.5 * torch.randn(B, H * G, W, D, generator=gen, device="cuda")).to(dtype) for _ in range(2)]
    n0 = _engine.launch_count()
    votes = _engine.snapkv_votes(list(zip(keys, qs)), W)
    assert _engine.launch_count() - n0 == 1
    for k, q, v in zip(keys, qs, votes):
        want = vote_reference(k, q, W)
        assert v.shape == want.shape and v.dtype == dtype
        got = v.float()
        ulp = 2.0 ** -8 if dtype == torch.bfloat16 else 2.0 ** -11
        # fp32 accumulate + ex2.approx, then ONE rounding to the cache dtype
        assert torch.all((got - want).abs() <= 2.5 * ulp * want + 1e-7), float(((got - want).abs() / (want + 1e-12)).max())
        # every window query distributes mass 1 over the keys it sees: votes sum to <= G*W per head
        assert torch.all(want.sum(-1) <= G * W + 1e-3)


def lse_reference(keys, queries, window):
    """fp32 log-sum-exp of every window query's attention row (what a flash-attention forward returns)."""
    B, H, S, D = keys.shape
    G = queries.size(1) // H
    P = S - window
    k = keys.float().repeat_interleave(G, dim=1)
    s = torch.matmul(queries.float(), k.transpose(-1, -2)) / math.sqrt(D)
    pos_q = P + torch.arange(window, device=keys.device).view(1, 1, window, 1)
    pos_k = torch.arange(S, device=keys.device).view(1, 1, 1, S)
    return torch.logsumexp(s.masked_fill(pos_k > pos_q, float("-inf")), dim=-1)   # [B, H*G, W]


@pytest.mark.parametrize("B,H,G,W,S,D,dtype", CASES[:4] + CASES[5:], ids=[f"B{c[0]}H{c[1]}G{c[2]}W{c[3]}S{c[4]}D{c[5]}" for c in CASES[:4] + CASES[5:]])
def test_single_pass_votes_with_caller_lse(B, H, G, W, S, D, dtype):
    """obs_lse: the kernel skips its first pass (the softmax denominators) and reads K once."""
    gen = torch.Generator(device="cuda").manual_seed(7 * S + D)
    keys = [torch.randn(B, H, S, D, generator=gen, device="cuda").to(dtype) for _ in range(2)]
    qs = [(1.5 * torch.randn(B, H * G, W, D, generator=gen, device="cuda")).to(dtype) for _ in range(2)]
    lse = [lse_reference(k, q, W) for k, q in zip(keys, qs)]
    votes = _engine.snapkv_votes(list(zip(keys, qs)), W, lse=lse)
    two_pass = _engine.snapkv_votes(list(zip(keys, qs)), W)
    ulp = 2.0 ** -8 if dtype == torch.bfloat16 else 2.0 ** -11
    for k, q, v, u in zip(keys, qs, votes, two_pass):
        want = vote_reference(k, q, W)
        got = v.float()
        assert torch.all((got - want).abs() <= 2.5 * ulp * want + 1e-7), float(((got - want).abs() / (want + 1e-12)).max())
        assert torch.all((got - u.float()).abs() <= 2.02 * ulp * u.float() + 1e-7)


def test_votes_at_the_c4_shape():
    """BASELINE configs[3] shape, one layer, one stream: S = 32768, 8 KV heads, 4 query heads per group, W = 32."""
    B, H, G, W, S, D = 1, 8, 4, 32, 32768, 128
    gen = torch.Generator(device="cuda").manual_seed(4)
    k = torch.randn(B, H, S, D, generator=gen, device="cuda").bfloat16()
    q = (1.5 * torch.randn(B, H * G, W, D, generator=gen, device="cuda")).bfloat16()
    v = _engine.snapkv_votes([(k, q)], W)[0]
    want = vote_reference(k, q, W)
    got = v.float()
    ulp = 2.0 ** -8
    assert torch.all((got - want).abs() <= 2.5 * ulp * want + 1e-7), float(((got - want).abs() / (want + 1e-12)).max())
    one = _engine.snapkv_votes([(k, q)], W, lse=[lse_reference(k, q, W)])[0]
    assert torch.all((one.float() - want).abs() <= 2.5 * ulp * want + 1e-7)


def test_strided_keys_and_queries():
    B, H, G, W, S, D = 1, 2, 4, 32, 600, 128
    k = torch.randn(B, S, H, D, device="cuda").bfloat16().permute(0, 2, 1, 3)      # [B,S,H,D] storage
    q = torch.randn(B, W, H * G, D, device="cuda").bfloat16().permute(0, 2, 1, 3)
    a = _engine.snapkv_votes([(k, q)], W)[0]
    b = _engine.snapkv_votes([(k.contiguous(), q.contiguous())], W)[0]
    assert torch.equal(a, b)


def test_snapkv_vote_mode_selects_the_highest_pooled_votes():
    torch.manual_seed(0)
    L, B, H, G, W, S, D, keep, pk = 3, 2, 2, 4, 32, 1500, 128, 256, 5
    kv = [(torch.randn(B, H, S, D, device="cuda").bfloat16(), torch.randn(B, H, S, D, device="cuda").bfloat16())
          for _ in range(L)]
    qs = [(2.0 * torch.randn(B, H * G, W, D, device="cuda")).bfloat16() for _ in range(L)]
    n0 = _engine.launch_count()
    out = kvcompress.snapkv_lite_compress(kv, observation_window=W, keep_size=keep, pooling_kernel=pk,
                                          skip_layers=[0], obs_queries=qs)
    assert _engine.launch_count() - n0 == 1  # vote, pool, select and gather of every layer in ONE launch
    assert out[0][0] is kv[0][0]
    for li in (1, 2):
        k_in, v_in = kv[li]
        k_out, v_out = out[li]
        assert k_out.shape == (B, H, keep, D)
        # window rows are the last W rows
        assert torch.equal(k_out[:, :, -W:], k_in[:, :, -W:]) and torch.equal(v_out[:, :, -W:], v_in[:, :, -W:])
        # recover the kept prefix rows by matching V rows (random data: rows are unique)
        votes = _engine.snapkv_votes([(k_in, qs[li])], W)[0]
        pooled = torch.nn.functional.avg_pool1d(votes.float().reshape(B * H, 1, -1), pk, 1, pk // 2).reshape(B, H, -1)
        pooled = pooled.to(torch.bfloat16).float()
        want_idx = torch.topk(pooled, keep - W, dim=-1)[1].sort(dim=-1)[0]
        rows_ref = torch.gather(v_in, 2, want_idx.unsqueeze(-1).expand(-1, -1, -1, D))
        # identical up to ties at the threshold: compare the multiset of pooled scores of the kept rows
        got_rows = v_out[:, :, :keep - W]
        same = (got_rows == rows_ref).all(-1).float().mean()
        assert same > 0.97, float(same)
        # out-of-place K rows correspond to the same positions as V rows
        k_ref = torch.gather(k_in, 2, want_idx.unsqueeze(-1).expand(-1, -1, -1, D))
        assert ((k_out[:, :, :keep - W] == k_ref).all(-1) == (got_rows == rows_ref).all(-1)).all()


@pytest.mark.parametrize("B,H,G,W,S,D,keep,pk,dtype", [
    (2, 2, 4, 32, 1500, 128, 256, 5, torch.bfloat16),
    (1, 3, 1, 32, 5000, 80, 512, 5, torch.bfloat16),
    (2, 2, 2, 16, 700, 64, 300, 3, torch.float16),
    (1, 2, 4, 32, 300, 128, 512, 5, torch.bfloat16),     # keep_size > S: untouched
    (1, 2, 4, 32, 600, 128, 512, 7, torch.bfloat16),
    (1, 2, 4, 32, 560, 128, 512, 1, torch.bfloat16),     # no pooling
    (1, 8, 4, 32, 32768, 128, 512, 5, torch.bfloat16),   # BASELINE configs[3] shape, one stream
])
def test_fused_vote_compress_equals_the_two_launch_form(B, H, G, W, S, D, keep, pk, dtype):
    """kvc_snapkv_vote_compress == kvc_snapkv_vote followed by kvc_compress_layers(GIVEN_SCORE): same votes (bit for
    bit), same kept rows, same K/V bytes; and the kept rows are the top-k of the pooled torch fp32 reference votes up
    to ties at the threshold."""
    from dataclasses import replace

    from kvcompress import _planner

    L = 2
    gen = torch.Generator(device="cuda").manual_seed(S + keep)
    kv = [(torch.randn(B, H, S, D, generator=gen, device="cuda").to(dtype),
           torch.randn(B, H, S, D, generator=gen, device="cuda").to(dtype)) for _ in range(L)]
    qs = [(2.0 * torch.randn(B, H * G, W, D, generator=gen, device="cuda")).to(dtype) for _ in range(L)]
    plans = _planner.plan_snapkv([S] * L, W, keep, pk, [])
    if plans[0].kind != _planner.GATHER:
        out = kvcompress.snapkv_lite_compress(kv, observation_window=W, keep_size=keep, pooling_kernel=pk, obs_queries=qs)
        assert all(a[0] is b[0] and a[1] is b[1] for a, b in zip(out, kv))
        return
    plans = [replace(p, score=_planner.SCORE_GIVEN_SCORE) for p in plans]
    n0 = _engine.launch_count()
    out, idx, votes = _engine.snapkv_vote_compress(kv, plans, qs, W, return_indices=True, return_votes=True)
    assert _engine.launch_count() - n0 == 1
    votes2 = _engine.snapkv_votes([(k, q) for (k, _), q in zip(kv, qs)], W)
    out2, idx2 = _engine.run_plans(kv, plans, given_scores=dict(enumerate(votes2)), return_indices=True)
    for li in range(L):
        assert torch.equal(votes[li], votes2[li])
        assert torch.equal(idx[li], idx2[li])
        assert torch.equal(out[li][0], out2[li][0]) and torch.equal(out[li][1], out2[li][1])
        ii = idx[li].long()
        assert torch.equal(out[li][0], torch.gather(kv[li][0], 2, ii.unsqueeze(-1).expand(-1, -1, -1, D)))
        assert torch.equal(out[li][1], torch.gather(kv[li][1], 2, ii.unsqueeze(-1).expand(-1, -1, -1, D)))
        assert torch.all(ii[..., 1:] > ii[..., :-1]) and torch.equal(ii[..., -W:], torch.arange(S - W, S, device="cuda").expand(B, H, W))
    api = kvcompress.snapkv_lite_compress(kv, observation_window=W, keep_size=keep, pooling_kernel=pk, obs_queries=qs)
    assert all(torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) for a, b in zip(api, out))


def test_prefix_beyond_the_fused_tail_takes_two_launches():
    """The fused tail keeps the prefix's radix keys in the (dead) key ring: ~90K rows at head_dim 64.  Longer prefixes
    run as vote launch + select/gather launch (workspace), same result contract."""
    B, H, G, W, S, D, keep = 1, 1, 2, 32, 100000, 64, 512
    gen = torch.Generator(device="cuda").manual_seed(8)
    k = torch.randn(B, H, S, D, generator=gen, device="cuda").bfloat16()
    v = torch.randn(B, H, S, D, generator=gen, device="cuda").bfloat16()
    q = (1.5 * torch.randn(B, H * G, W, D, generator=gen, device="cuda")).bfloat16()
    n0 = _engine.launch_count()
    out = kvcompress.snapkv_lite_compress([(k, v)], observation_window=W, keep_size=keep, obs_queries=[q])
    assert _engine.launch_count() - n0 == 2
    votes = _engine.snapkv_votes([(k, q)], W)[0]
    want, idx = _engine.run_plans([(k, v)], [_planner_plan(S, W, keep)], given_scores={0: votes}, return_indices=True)
    assert torch.equal(out[0][0], want[0][0]) and torch.equal(out[0][1], want[0][1])
    assert torch.equal(out[0][1], torch.gather(v, 2, idx[0].long().unsqueeze(-1).expand(-1, -1, -1, D)))


def _planner_plan(S, W, keep, pk=5):
    from kvcompress import _planner

    return _planner.LayerPlan(_planner.GATHER, S, 0, 0, S - W, keep - W, W, _planner.SCORE_GIVEN_SCORE, pk)


def test_vote_errors():
    k = torch.randn(1, 2, 300, 128, device="cuda")
    q = torch.randn(1, 8, 32, 128, device="cuda")
    with pytest.raises(ValueError, match="bfloat16/float16"):
        _engine.snapkv_votes([(k, q)], 32)
    with pytest.raises(ValueError, match="exceeds the 128 query rows"):
        _engine.snapkv_votes([(k.bfloat16(), torch.randn(1, 16, 32, 128, device="cuda").bfloat16())], 32)
    with pytest.raises(ValueError, match="obs_queries"):
        kvcompress.snapkv_lite_compress([(k.bfloat16(), k.bfloat16())], obs_queries=[])


def test_vote_mode_in_place_on_a_slab_keeps_the_rows_of_the_function():
    """``KVSlabCache.compress_("snapkv_lite", obs_queries=...)``: vote launch + in-place compaction with the votes as
    caller-supplied scores (``KVC_SCORE_GIVEN_SCORE`` in ``kvc_slab_compress``) against the fused function call."""
    from kvcompress import KVSlabCache

    torch.manual_seed(5)
    L, B, H, G, W, S, D, keep = 3, 2, 2, 4, 32, 1500, 128, 256
    kv = [(torch.randn(B, H, S, D, device="cuda").bfloat16(), torch.randn(B, H, S, D, device="cuda").bfloat16())
          for _ in range(L)]
    qs = [(2.0 * torch.randn(B, H * G, W, D, device="cuda")).bfloat16() for _ in range(L)]
    for pk, skip in ((5, []), (4, [1]), (1, [])):
        kw = dict(observation_window=W, keep_size=keep, pooling_kernel=pk, skip_layers=skip)
        want = kvcompress.snapkv_lite_compress(kv, obs_queries=qs, **kw)
        slab = KVSlabCache.from_legacy_cache(kv, capacity=S + 8)
        n0 = _engine.launch_count()
        slab.compress_("snapkv_lite", obs_queries=qs, **kw)
        assert _engine.launch_count() - n0 == 2          # the vote, then the in-place select + slide
        assert slab.lengths == [k.size(2) for k, _ in want]
        for li in range(L):
            assert torch.equal(slab[li][0], want[li][0]) and torch.equal(slab[li][1], want[li][1]), (pk, li)
    # single pass with the caller's log-sum-exp
    lse = [lse_reference(k, q, W) for (k, _), q in zip(kv, qs)]
    want = kvcompress.snapkv_lite_compress(kv, obs_queries=qs, obs_lse=lse, observation_window=W, keep_size=keep)
    slab = KVSlabCache.from_legacy_cache(kv, capacity=S + 8)
    slab.compress_("snapkv_lite", obs_queries=qs, obs_lse=lse, observation_window=W, keep_size=keep)
    for li in range(L):
        assert torch.equal(slab[li][0], want[li][0]) and torch.equal(slab[li][1], want[li][1])
