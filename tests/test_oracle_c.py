"""Cross-check the C port (oracle/kvc_oracle.c, the CPU baseline) against the numpy oracle."""

import numpy as np
import pytest

import cases
from oracle import kvc_oracle as O
from oracle import kvc_oracle_c as OC

SUBSET = [c for c in cases.all_cases()
          if c["method"] != "recent_only" and not c["name"].startswith(("c3/", "c4/", "c5/"))]


@pytest.mark.parametrize("case", SUBSET, ids=[c["name"] for c in SUBSET])
def test_c_port_matches_numpy_oracle(case):
    layers = cases.case_cache(case)
    dtype = case["dtype"]
    want = O.METHODS[case["method"]](layers, dtype, **case["kwargs"])
    outs, got, _ = OC.run_method(case["method"], layers, dtype, nthreads=2, **case["kwargs"])
    for li, (w, g) in enumerate(zip(want, got)):
        assert w.untouched == g.untouched and w.is_view == g.is_view and w.out_len == g.out_len
        if w.untouched:
            assert outs[li][0] is layers[li][0]
            continue
        if w.is_view or w.k_sel == 0:
            assert np.array_equal(outs[li][0], O.take_rows(layers[li][0], w.rows))
            assert np.array_equal(outs[li][1], O.take_rows(layers[li][1], w.rows))
            continue
        info = O.check_layer(layers[li][0], dtype, w, g.rows)
        assert info["valid"], (case["name"], li, info)
        # both are stable sorts over norms that agree except on fp32 rounding boundaries
        assert info["identical_heads"] >= info["heads"] - 1, (case["name"], li, info)
        assert np.array_equal(outs[li][0], O.take_rows(layers[li][0], g.rows))
        assert np.array_equal(outs[li][1], O.take_rows(layers[li][1], g.rows))


@pytest.mark.parametrize("dtype", ["f32", "f16", "bf16"])
def test_c_norms_and_conversions(dtype):
    layers = cases.make_cache(5, [400], 2, 2, 80, dtype, "spread")
    K = layers[0][0]
    got = OC.norms(K, dtype, 3, 390)
    want = O.key_norms(K[:, :, 3:390], dtype)
    if dtype == "f32":
        np.testing.assert_allclose(got, want, rtol=1e-6)
    else:
        assert np.mean(got == want) > 0.999
    # exhaustive 16-bit round trips through the C conversion helpers: the norm of a one-hot row is |x|
    # (x*x is exact in fp32 for 11-/8-bit significands; bf16 restricted to exponents whose square fits fp32)
    if dtype != "f32":
        bits = np.arange(0, 0x7C00, dtype=np.uint16) if dtype == "f16" else np.arange(0x2000, 0x5F80, dtype=np.uint16)
        one_hot = np.zeros((1, 1, bits.size, 8), dtype=np.uint16)
        one_hot[..., 0] = bits
        arr = one_hot.view(np.float16) if dtype == "f16" else one_hot
        n = OC.norms(arr, dtype, 0, bits.size)[0, 0]
        np.testing.assert_array_equal(n, O.to_f32(arr[0, 0, :, 0], dtype))


def test_thread_count_is_reported():
    assert OC.max_threads() >= 1
