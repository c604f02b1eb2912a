"""CPU oracle for the kvcompress compression hot path — TEST INFRASTRUCTURE, NOT PRODUCT.

A numpy restatement of what the reference's eight compress functions compute
(/root/reference/kvcompress/methods/*.py; file:line cited at each function), written in
*index space*: every method yields, per layer, either "untouched" or an int64 array
``rows[B, H, C]`` of the kept token positions in ascending order; ``take_rows`` then gathers
K and V.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg may
import this package; the product (``cs3602-llm-inference-acceleration_b200/``) never does.

Pinning: the reference's own tests hold no golden vectors for this path, so the
oracle is pinned against outputs of the reference itself, run in the build container by
``tests/golden/make_golden.py`` and committed as ``tests/golden/*.npz``
(``tests/test_oracle_golden.py`` checks every one of them).

Numerics the reference inherits from torch and that are restated here:
  * ``torch.norm(K, p=2, dim=-1)`` accumulates in fp32 and returns the INPUT dtype; here the norm
    is computed in float64, rounded to fp32, then to the input dtype;
  * ``argsort`` / ``topk`` tie order is unspecified in torch; the oracle (and the CUDA path) define
    it as *lowest token index first* (a stable sort), and ``selection_is_valid`` accepts any
    selection that differs from that only among keys that are equal — or that would be equal under
    a 1e-6 relative perturbation of the fp32 norm (BASELINE.json north_star);
  * ``avg_pool1d(k, stride 1, pad k//2)`` = fp32 left-to-right sum of the zero-padded window
    divided by k, rounded to the input dtype (count_include_pad=True).

Array conventions: fp32 caches are ``np.float32`` arrays, fp16 caches ``np.float16``, bf16 caches
are ``np.uint16`` arrays of bit patterns with ``dtype="bf16"``.
"""

from __future__ import annotations

from math import ceil
from typing import List, Optional, Sequence, Tuple

import numpy as np

F32, F16, BF16 = "f32", "f16", "bf16"
REL_TOL = 1e-6  # north_star: fp32 norms within 1e-6 relative


# ----------------------------------------------------------------------------- dtype helpers
def bf16_bits_from_f32(x: np.ndarray) -> np.ndarray:
    """float32 -> bfloat16 bit patterns, round-to-nearest-even (what ``tensor.to(torch.bfloat16)`` does)."""
    bits = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    rounded = (bits + (np.uint32(0x7FFF) + ((bits >> np.uint32(16)) & np.uint32(1)))) >> np.uint32(16)
    return rounded.astype(np.uint16)


def f32_from_bf16_bits(bits: np.ndarray) -> np.ndarray:
    return (np.ascontiguousarray(bits, dtype=np.uint16).astype(np.uint32) << np.uint32(16)).view(np.float32)


def to_f32(a: np.ndarray, dtype: str) -> np.ndarray:
    """Exact float32 values of a stored array."""
    if dtype == BF16:
        return f32_from_bf16_bits(a)
    return np.asarray(a).astype(np.float32)


def round_to(x: np.ndarray, dtype: str) -> np.ndarray:
    """Round float32 values to the storage dtype; result as float32 (exact)."""
    x = np.asarray(x, dtype=np.float32)
    if dtype == F32:
        return x
    if dtype == F16:
        return x.astype(np.float16).astype(np.float32)
    return f32_from_bf16_bits(bf16_bits_from_f32(x))


def store(x: np.ndarray, dtype: str) -> np.ndarray:
    """float32 values -> storage array of `dtype`."""
    x = np.asarray(x, dtype=np.float32)
    if dtype == F32:
        return x
    if dtype == F16:
        return x.astype(np.float16)
    return bf16_bits_from_f32(x)


# ----------------------------------------------------------------------------- norms / keys
def exact_norms(K: np.ndarray, dtype: str) -> np.ndarray:
    """float64 ||K[b,h,s,:]||_2 of the stored values."""
    k64 = to_f32(K, dtype).astype(np.float64)
    return np.sqrt(np.einsum("bhsd,bhsd->bhs", k64, k64))


def key_norms(K: np.ndarray, dtype: str) -> np.ndarray:
    """``torch.norm(K, p=2, dim=-1)``: value as float32, already rounded to the input dtype
    (e.g. l2_compress.py:70)."""
    return round_to(exact_norms(K, dtype).astype(np.float32), dtype)


def key_norm_interval(K: np.ndarray, dtype: str, rel_tol: float = REL_TOL) -> Tuple[np.ndarray, np.ndarray]:
    """Lowest / highest key value an implementation may legitimately see for each token: the exact
    norm perturbed by +-rel_tol, rounded to fp32 and then to the input dtype."""
    n = exact_norms(K, dtype)
    lo = round_to((n * (1.0 - rel_tol)).astype(np.float32), dtype)
    hi = round_to((n * (1.0 + rel_tol)).astype(np.float32), dtype)
    return lo, hi


def _pool_sum(scores: np.ndarray, kernel: int) -> np.ndarray:
    """fp32 left-to-right sum over the zero-padded window [i - kernel//2, i - kernel//2 + kernel)."""
    P = scores.shape[-1]
    pad = kernel // 2
    padded = np.zeros(scores.shape[:-1] + (P + 2 * pad + kernel,), dtype=np.float32)
    padded[..., pad:pad + P] = scores
    acc = np.zeros_like(scores, dtype=np.float32)
    for t in range(kernel):
        acc = (acc + padded[..., t:t + P]).astype(np.float32)
    return acc


def snapkv_scores(norms: np.ndarray, dtype: str, pooling_kernel: int, max_norm: Optional[np.ndarray] = None) -> np.ndarray:
    """Importance of prefix tokens, snapkv_lite.py:96-121: ``(max + 1e-6) - norm`` in the input dtype,
    then ``avg_pool1d(kernel, stride 1, pad kernel//2)`` iff ``kernel > 1 and P >= kernel``."""
    P = norms.shape[-1]
    mx = norms.max(axis=-1, keepdims=True) if max_norm is None else max_norm
    mxe = round_to((mx.astype(np.float32) + np.float32(1e-6)).astype(np.float32), dtype)
    scores = round_to((mxe - norms).astype(np.float32), dtype)
    if pooling_kernel > 1 and P >= pooling_kernel:
        pooled = (_pool_sum(scores, pooling_kernel) / np.float32(pooling_kernel)).astype(np.float32)
        scores = round_to(pooled, dtype)
    return scores


# ----------------------------------------------------------------------------- selection
def lowest_k(keys: np.ndarray, k: int) -> np.ndarray:
    """``keys.argsort()[..., :k]`` then ``sort``: positions of the k smallest keys, ascending;
    ties -> lowest index (e.g. h2o_l2.py:128-132)."""
    order = np.argsort(keys, axis=-1, kind="stable")[..., :k]
    return np.sort(order, axis=-1).astype(np.int64)


def highest_k(keys: np.ndarray, k: int) -> np.ndarray:
    """``argsort(descending=True)[..., :k]`` / ``topk(k)`` then ``sort``; ties -> lowest index
    (fix_size_l2.py:110-114, snapkv_lite.py:134-137)."""
    order = np.argsort(-keys.astype(np.float64), axis=-1, kind="stable")[..., :k]
    return np.sort(order, axis=-1).astype(np.int64)


def selection_is_valid(sel: np.ndarray, lo: np.ndarray, hi: np.ndarray, largest: bool = False) -> np.ndarray:
    """Tie-aware acceptance of a selection (SURVEY.md §8c rule 2).

    sel: [..., k] chosen positions; lo/hi: [..., R] key interval per position.  A selection of the
    k smallest is valid iff some assignment key_i in [lo_i, hi_i] makes it a correct answer with
    arbitrary tie-breaking:  max(lo[selected]) <= min(hi[not selected])  — and the positions are
    strictly ascending and in range.  Returns a boolean array over the leading dims."""
    sel = np.asarray(sel)
    R = lo.shape[-1]
    lead = lo.shape[:-1]
    flat_sel = sel.reshape(-1, sel.shape[-1])
    flat_lo = lo.reshape(-1, R)
    flat_hi = hi.reshape(-1, R)
    ok = np.ones(flat_sel.shape[0], dtype=bool)
    for r in range(flat_sel.shape[0]):
        s = flat_sel[r]
        if s.size and (s.min() < 0 or s.max() >= R or np.any(np.diff(s) <= 0)):
            ok[r] = False
            continue
        mask = np.zeros(R, dtype=bool)
        mask[s] = True
        if s.size == 0 or s.size == R:
            continue
        if largest:
            ok[r] = flat_hi[r][mask].min() >= flat_lo[r][~mask].max()
        else:
            ok[r] = flat_lo[r][mask].max() <= flat_hi[r][~mask].min()
    return ok.reshape(lead)


# ----------------------------------------------------------------------------- row builders
def _head(S: int, n: int) -> np.ndarray:
    return np.arange(S)[:n]  # x[:, :, :n]


def _last(S: int, n: int) -> np.ndarray:
    return np.arange(S)[-n:]  # x[:, :, -n:]   (-0: is everything)


_NONE = np.zeros(0, dtype=np.int64)


class LayerResult:
    """What a method does to one layer: untouched, or   head rows + selected rows + tail rows.

    `head` / `tail` are 1-D row lists shared by every (b,h) (a prefix and a suffix of the layer);
    `sel` is the [B,H,k_sel] selection in absolute rows (None while deferred); `is_view` marks the
    paths where the reference returns a slice view instead of a fresh tensor."""

    def __init__(self, shape=None, head=_NONE, tail=_NONE, sel=None, region=(0, 0), k_sel=0, mode="none",
                 pool_kernel=1, is_view=False, untouched=False):
        self.untouched = untouched
        self.shape = shape          # (B, H, S, D)
        self.head = np.asarray(head, dtype=np.int64)
        self.tail = np.asarray(tail, dtype=np.int64)
        self.sel = sel
        self.region = region        # [lo, hi) the selection runs over
        self.k_sel = k_sel
        self.mode = mode            # "none" | "low" | "high" | "snapkv" | "random"
        self.pool_kernel = pool_kernel
        self.is_view = is_view

    @property
    def out_len(self) -> int:
        return self.shape[2] if self.untouched else len(self.head) + self.k_sel + len(self.tail)

    @property
    def rows(self) -> Optional[np.ndarray]:
        """int64 [B, H, C] kept rows, ascending by construction."""
        if self.untouched:
            return None
        B, H = self.shape[0], self.shape[1]
        parts = [np.broadcast_to(self.head, (B, H, len(self.head)))]
        if self.k_sel > 0:
            if self.sel is None:
                raise RuntimeError("selection deferred: run it with a backend first")
            parts.append(np.asarray(self.sel, dtype=np.int64))
        parts.append(np.broadcast_to(self.tail, (B, H, len(self.tail))))
        return np.concatenate(parts, axis=-1)


def _same(K) -> LayerResult:
    return LayerResult(shape=K.shape, untouched=True)


def numpy_select(K: np.ndarray, dtype: str, lo: int, hi: int, k: int, mode: str, pool_kernel: int = 1) -> np.ndarray:
    """norm -> (score) -> k best of rows [lo, hi), ascending absolute rows [B,H,k]."""
    norms = key_norms(K[:, :, lo:hi], dtype)
    if mode == "low":
        return lowest_k(norms, k) + lo
    if mode == "high":
        return highest_k(norms, k) + lo
    if mode == "snapkv":
        return highest_k(snapkv_scores(norms, dtype, pool_kernel), k) + lo
    raise ValueError(mode)


def _selected(K, dtype, head, lo, hi, k, tail, mode, select, pool_kernel=1) -> LayerResult:
    sel = None
    if k > 0 and select is not None:
        sel = select(K, dtype, lo, hi, k, mode, pool_kernel)
    return LayerResult(shape=K.shape, head=head, tail=tail, sel=sel, region=(lo, hi), k_sel=k,
                       mode=mode if k > 0 else "none", pool_kernel=pool_kernel)


def _shape(K: np.ndarray):
    return K.shape[0], K.shape[1], K.shape[2], K.shape[3]


# ----------------------------------------------------------------------------- the eight methods
# Every method takes `select=`: the routine that ranks the selection region (default: numpy above).
# `select=None` defers the ranking and returns only the per-layer descriptors (used to drive the C
# port in oracle/kvc_oracle.c and to compute output lengths without touching the data).
def l2_compress(layers, dtype, keep_ratio=1.0, prune_after=1000, skip_layers=(0, 1), select=numpy_select):
    """reference methods/l2_compress.py:18-92."""
    out = []
    for li, (K, _V) in enumerate(layers):
        S = K.shape[2]
        keep = ceil(keep_ratio * S)
        if keep_ratio >= 1.0 or S <= prune_after or li in skip_layers or keep >= S:  # :48-65
            out.append(_same(K))
            continue
        out.append(_selected(K, dtype, _NONE, 0, S, keep, _NONE, "low", select))  # :70-88
    return out


def fix_size_l2_compress(layers, dtype, fix_kv_size=1024, keep_ratio=0.0, strategy="keep_low", skip_layers=(0, 1),
                         random_rows=None, select=numpy_select):
    """reference methods/fix_size_l2.py:15-154.  For strategy="random" the caller passes the rows the
    torch generator produced (``random_rows[layer] = [B,H,k]``) — RNG streams are torch's."""
    out = []
    for li, (K, _V) in enumerate(layers):
        S = K.shape[2]
        if S <= fix_kv_size or li in skip_layers:  # :69-74
            out.append(_same(K))
            continue
        protected = min(int(fix_kv_size * keep_ratio), S)  # :79-80
        zone_end = S - protected
        budget = fix_kv_size - protected
        if budget <= 0:  # :88-93
            out.append(LayerResult(shape=K.shape, tail=_last(S, protected), is_view=True))
            continue
        if zone_end <= budget:  # :95-97
            out.append(_same(K))
            continue
        tail = _last(S, protected) if protected > 0 else _NONE  # :141-150
        if strategy == "keep_low":  # :104-108
            out.append(_selected(K, dtype, _NONE, 0, zone_end, budget, tail, "low", select))
        elif strategy == "keep_high":  # :110-114
            out.append(_selected(K, dtype, _NONE, 0, zone_end, budget, tail, "high", select))
        elif strategy == "random":  # :116-124, :129
            rows = np.sort(np.asarray(random_rows[li], dtype=np.int64), axis=-1)
            out.append(LayerResult(shape=K.shape, tail=tail, sel=rows, region=(0, zone_end), k_sel=budget, mode="random"))
        else:
            raise ValueError(f"Unknown strategy: {strategy}")  # :126
    return out


def streaming_llm_compress(layers, dtype, start_size=4, recent_size=508, skip_layers=(), select=None):
    """reference methods/streaming_llm.py:19-111."""
    out = []
    for li, (K, _V) in enumerate(layers):
        S = K.shape[2]
        if S <= start_size + recent_size or li in skip_layers:  # :88-93
            out.append(_same(K))
            continue
        out.append(LayerResult(shape=K.shape, head=_head(S, start_size), tail=_last(S, recent_size)))  # :99-107
    return out


def evict_for_space(layers, dtype, num_coming, start_size=4, recent_size=508, skip_layers=(), select=None):
    """reference methods/streaming_llm.py:114-170."""
    out = []
    for li, (K, _V) in enumerate(layers):
        S = K.shape[2]
        if S + num_coming <= start_size + recent_size or li in skip_layers:  # :147-152
            out.append(_same(K))
            continue
        recent = recent_size - num_coming  # :155-157
        if recent <= 0:
            recent = recent_size
        out.append(LayerResult(shape=K.shape, head=_head(S, start_size), tail=_last(S, recent)))
    return out


def recent_only_compress(layers, dtype, window_size=512, skip_layers=(0, 1), select=None):
    """reference methods/recent_only.py:16-70 (returns views)."""
    out = []
    for li, (K, _V) in enumerate(layers):
        S = K.shape[2]
        if S <= window_size or li in skip_layers:  # :57-62
            out.append(_same(K))
            continue
        out.append(LayerResult(shape=K.shape, tail=_last(S, window_size), is_view=True))  # :65-66
    return out


def h2o_l2_compress(layers, dtype, start_size=4, heavy_hitter_size=64, recent_size=444, skip_layers=(),
                    select=numpy_select):
    """reference methods/h2o_l2.py:25-153."""
    out = []
    for li, (K, _V) in enumerate(layers):
        S = K.shape[2]
        if S <= start_size + heavy_hitter_size + recent_size or li in skip_layers:  # :81-86
            out.append(_same(K))
            continue
        lo, hi = start_size, S - recent_size  # :95-96
        if hi <= lo:  # :99-109
            out.append(LayerResult(shape=K.shape, head=_head(S, start_size), tail=_last(S, recent_size)))
            continue
        k = min(heavy_hitter_size, hi - lo)  # :125
        out.append(_selected(K, dtype, _head(S, start_size), lo, hi, k, _last(S, recent_size), "low", select))
    return out


def snapkv_lite_compress(layers, dtype, observation_window=32, keep_size=512, pooling_kernel=5, skip_layers=(),
                         select=numpy_select):
    """reference methods/snapkv_lite.py:24-154."""
    out = []
    for li, (K, _V) in enumerate(layers):
        S = K.shape[2]
        P = S - observation_window  # :83
        if S <= keep_size or li in skip_layers or P <= 0:  # :70-86
            out.append(_same(K))
            continue
        k = min(keep_size - observation_window, P)  # :125-126
        if k <= 0:  # :128-131
            out.append(LayerResult(shape=K.shape, tail=_last(S, observation_window), is_view=True))
            continue
        out.append(_selected(K, dtype, _NONE, 0, P, k, _last(S, observation_window), "snapkv", select,
                             pool_kernel=pooling_kernel))  # :96-150
    return out


def pyramid_layer_sizes(num_layers, base_size=512, layer_decay=0.9, min_size=64, profile="exponential") -> List[int]:
    """reference methods/pyramid_kv.py:84-97."""
    sizes = []
    for i in range(num_layers):
        if profile == "exponential":
            size = int(base_size * (layer_decay ** i))
        elif profile == "linear":
            size = int(base_size - i * ((base_size - min_size) / max(num_layers - 1, 1)))
        else:
            size = base_size
        sizes.append(max(size, min_size))
    return sizes


def _sinks_middle_recent(K, dtype, target, start, select) -> LayerResult:
    """pyramid_kv.py:115-181 and the hard branch of adaptive_l2.py:86-143 share this shape."""
    S = K.shape[2]
    recent = target // 2
    middle_budget = target - start - recent
    if middle_budget <= 0:  # pyramid :119-124, adaptive :93-98
        return LayerResult(shape=K.shape, tail=_last(S, target), is_view=True)
    lo, hi = start, S - recent
    if hi <= lo:  # pyramid :130-140, adaptive :104-113
        return LayerResult(shape=K.shape, head=_head(S, start), tail=_last(S, target - start))
    k = min(middle_budget, hi - lo)
    return _selected(K, dtype, _head(S, start), lo, hi, k, _last(S, recent), "low", select)


def pyramid_kv_compress(layers, dtype, base_size=512, layer_decay=0.9, min_size=64, profile="exponential",
                        skip_layers=(), select=numpy_select):
    """reference methods/pyramid_kv.py:26-185."""
    sizes = pyramid_layer_sizes(len(layers), base_size, layer_decay, min_size, profile)
    out = []
    for li, (K, _V) in enumerate(layers):
        S = K.shape[2]
        target = sizes[li]
        if S <= target or li in skip_layers:  # :104-109
            out.append(_same(K))
            continue
        out.append(_sinks_middle_recent(K, dtype, target, min(4, target // 8), select))
    return out


def adaptive_l2_compress(layers, dtype, target_size=512, soft_limit=256, hard_limit=1024, keep_ratio_min=0.3,
                         keep_ratio_max=0.9, skip_layers=(), select=numpy_select):
    """reference methods/adaptive_l2.py:20-201."""
    out = []
    for li, (K, _V) in enumerate(layers):
        S = K.shape[2]
        if li in skip_layers or S <= soft_limit:  # :71-77
            out.append(_same(K))
            continue
        if S > hard_limit:  # :81-145
            out.append(_same(K) if S <= target_size else _sinks_middle_recent(K, dtype, target_size, 4, select))
            continue
        progress = (S - soft_limit) / (hard_limit - soft_limit)  # :150
        ratio = keep_ratio_max - progress * (keep_ratio_max - keep_ratio_min)  # :151
        n = max(int(S * ratio), soft_limit)  # :153-154
        if n >= S:
            out.append(_same(K))
            continue
        recent = int(n * 0.2)  # :160
        k = n - recent
        if k <= 0:  # :163-168
            out.append(LayerResult(shape=K.shape, tail=_last(S, n), is_view=True))
            continue
        hi = S - recent
        if hi <= k:  # :173-174
            out.append(_same(K))
            continue
        out.append(_selected(K, dtype, _NONE, 0, hi, k, _last(S, recent), "low", select))  # :176-197
    return out


METHODS = {
    "l2_compress": l2_compress,
    "fix_size_l2": fix_size_l2_compress,
    "streaming_llm": streaming_llm_compress,
    "recent_only": recent_only_compress,
    "h2o_l2": h2o_l2_compress,
    "snapkv_lite": snapkv_lite_compress,
    "pyramid_kv": pyramid_kv_compress,
    "adaptive_l2": adaptive_l2_compress,
}


# ----------------------------------------------------------------------------- gather + checks
def take_rows(X: np.ndarray, rows: np.ndarray) -> np.ndarray:
    """``torch.gather(X, 2, rows[..., None].expand(..., D))`` (e.g. l2_compress.py:82-88)."""
    return np.take_along_axis(X, rows[..., None], axis=2)


def apply(layers, results: Sequence[LayerResult]):
    """Materialise a method's result as a list of (K, V) numpy arrays."""
    out = []
    for (K, V), res in zip(layers, results):
        out.append((K, V) if res.untouched else (take_rows(K, res.rows), take_rows(V, res.rows)))
    return out


def out_lengths(layers, results: Sequence[LayerResult]) -> List[int]:
    return [r.out_len for r in results]


def algorithmic_bytes(results: Sequence[LayerResult], elem_bytes: int) -> int:
    """e*B*H*D*(R + 4*C) summed over layers that move data."""
    total = 0
    for r in results:
        if r.untouched or r.is_view:
            continue
        B, H, _S, D = r.shape
        region = (r.region[1] - r.region[0]) if (r.k_sel > 0 and r.mode != "random") else 0
        total += elem_bytes * B * H * D * (region + 4 * r.out_len)
    return total


def key_interval(K: np.ndarray, dtype: str, res: LayerResult, rel_tol: float = REL_TOL):
    """Per-position [lo, hi] key interval of the selection region for `selection_is_valid`."""
    lo, hi = res.region
    a, b = key_norm_interval(K[:, :, lo:hi], dtype, rel_tol)
    if res.mode == "snapkv":
        # scores fall when norms rise and rise with the max: propagate the interval end points
        s_lo = snapkv_scores(b, dtype, res.pool_kernel, max_norm=a.max(axis=-1, keepdims=True))
        s_hi = snapkv_scores(a, dtype, res.pool_kernel, max_norm=b.max(axis=-1, keepdims=True))
        return s_lo, s_hi
    return a, b


def check_layer(K: np.ndarray, dtype: str, res: LayerResult, got_rows: np.ndarray, rel_tol: float = REL_TOL) -> dict:
    """Grade an implementation's kept rows for one layer against this oracle.

    Returns {"valid": all heads pass the tie-aware rule and sinks/tail match exactly,
             "identical_heads": heads whose rows equal the oracle's exactly, "heads": B*H}."""
    want = res.rows
    got_rows = np.asarray(got_rows, dtype=np.int64)
    heads = want.shape[0] * want.shape[1]
    if got_rows.shape != want.shape:
        return {"valid": False, "identical_heads": 0, "heads": heads, "why": "shape"}
    same = np.all(got_rows == want, axis=-1)
    info = {"valid": True, "identical_heads": int(same.sum()), "heads": heads}
    if res.mode in ("none", "random") or res.k_sel == 0:
        info["valid"] = bool(same.all())
        return info
    lo, _hi = res.region
    n_head = len(res.head)
    fixed = np.ones(want.shape[-1], dtype=bool)
    fixed[n_head:n_head + res.k_sel] = False
    if not np.array_equal(got_rows[..., fixed], want[..., fixed]):
        info.update(valid=False, why="sink/tail rows differ")
        return info
    a, b = key_interval(K, dtype, res, rel_tol)
    sel = got_rows[..., n_head:n_head + res.k_sel] - lo
    ok = selection_is_valid(sel, a, b, largest=res.mode in ("high", "snapkv"))
    info["valid"] = bool(ok.all())
    if not info["valid"]:
        info["why"] = f"{int((~ok).sum())} heads violate the tie-aware selection rule"
    return info
