"""The C-ABI library loads and exports every function include/kvc.h declares (no compute here)."""

import ctypes
import os
import re

import pytest

from kvcompress import _engine

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "kvc.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kvc_[a-z_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = declared_functions()
    assert {"kvc_compress_layers", "kvc_key_norms", "kvc_select", "kvc_abi_version"} <= set(names)
    lib = ctypes.CDLL(_engine.library_path())
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/kvc.h but not exported"


def test_library_metadata_and_argument_checks():
    lib = _engine.load_library()
    assert lib.kvc_abi_version() == 5
    assert b"sm_100a" in lib.kvc_build_info()
    assert lib.kvc_status_string(0) == b"ok"
    assert lib.kvc_launch_count() >= 0
    # bf16 keys are 2 bytes: a 32K-token region fits on chip with room to spare; fp32 up to ~52K rows
    assert lib.kvc_max_region_rows(2, 512) >= 65536
    assert 49152 <= lib.kvc_max_region_rows(0, 512) < 65536
    # argument validation happens before any CUDA call, so it is testable without a GPU
    assert lib.kvc_compress_layers(None, 1, None, None, None) == 1
    shape = _engine._SHAPE.pack(1, 1, 12, 2, 0)  # D=12 bf16 -> 24-byte rows: not 16-byte aligned
    plan = _engine._PLAN.pack(10, 1, 0, 0, 0, 1, 0, 1)
    io = _engine._IO.pack(16, 16, 16, 16, 120, 120, 12, 120, 120, 12, 0, 0, 0, 0, 0, 0)
    assert lib.kvc_compress_layers(shape, 1, plan, io, None) == 2
    shape = _engine._SHAPE.pack(1, 1, 16, 2, 0)
    bad_plan = _engine._PLAN.pack(10, 1, 0, 20, 3, 1, 1, 1)  # sel_hi > seq_len
    assert lib.kvc_compress_layers(shape, 1, bad_plan, io, None) == 1


def test_structs_match_header_sizes():
    assert _engine._PLAN.size == 32 and _engine._IO.size == 128 and _engine._SHAPE.size == 20


def test_library_ignores_the_environment(monkeypatch):
    """The product build reads no environment variable (launch-shape overrides and stage isolation exist only in a
    -DKVC_LAB build): the planner's answers do not move when the old knobs are set."""
    lib = _engine.load_library()
    shape = _engine._SHAPE.pack(2, 8, 128, 2, 0)
    plan = _engine._PLAN.pack(200000, 4, 4, 199000, 64, 444, 1, 1)
    before = int(lib.kvc_workspace_bytes(shape, 1, plan))
    for name in ("KVC_TMA_NT", "KVC_TMA_CTAS", "KVC_TMA_NSW", "KVC_TMA_UPC", "KVC_VOTE_DEBUG", "KVC_FORCE_LDG"):
        monkeypatch.setenv(name, "1")
    assert int(lib.kvc_workspace_bytes(shape, 1, plan)) == before > 0
    text = open(_engine.library_path(), "rb").read()
    assert b"KVC_VOTE_DEBUG" not in text and b"KVC_TMA_NT" not in text


def test_vote_argument_checks():
    lib = _engine.load_library()
    shape32 = _engine._SHAPE.pack(2, 8, 128, 0, 0)        # fp32 caches: the vote runs on 16-bit data only
    layer = _engine._VOTE.pack(16, 16, 16, 1, 1, 128, 1, 1, 128, 4096, 0, 0)
    assert lib.kvc_snapkv_vote(shape32, 1, layer, 4, 32, None) == 2
    shape96 = _engine._SHAPE.pack(2, 8, 96, 2, 0)         # head_dim 96: no vote kernel for that row width
    assert lib.kvc_snapkv_vote(shape96, 1, layer, 4, 32, None) == 2
    shape = _engine._SHAPE.pack(2, 8, 128, 2, 0)
    assert lib.kvc_snapkv_vote(shape, 1, layer, 8, 32, None) == 2   # 256 query rows > one MMA
    assert lib.kvc_snapkv_vote(shape, 1, None, 4, 32, None) == 1


def test_integration_stub_uses_the_current_struct_layouts():
    """INTEGRATION.md section 3 shows a ctypes stub a maintainer would paste: its struct formats must be the binding's."""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    for fmt in (_engine._PLAN.format, _engine._IO.format, _engine._SHAPE.format):
        assert f'struct.Struct("{fmt}")' in text, fmt
    header = open(os.path.join(ROOT, "include", "kvc.h")).read()
    assert f"#define KVC_ABI_VERSION {_engine.KVC_ABI_VERSION}" in header


def test_launch_shapes_of_the_baseline_configs():
    """kvc_launch_shape is the planner's answer without a GPU: many-wave launches keep 256 threads x 3 resident CTAs,
    32K-row regions (64 KB of keys) take one 512-thread CTA per SM, and a launch of a few large units (c1: 960 units of
    2.75 MB) is ranked by its estimated length — one CTA per SM, a short last wave.  (A scoring slip once sent c5 to
    256 x 2 and cost it 7 %.)"""
    import ctypes as C

    from kvcompress import _planner as P

    lib = _engine.load_library()
    lib.kvc_launch_shape.restype = C.c_int

    def shape_of(B, H, D, dtype, plans):
        rec = b"".join(_engine._PLAN.pack(p.seq_len, p.sink, p.sel_lo, p.sel_hi, p.k_sel, p.tail, p.score, p.pool_kernel)
                       for p in plans if p.kind == P.GATHER)
        n = sum(p.kind == P.GATHER for p in plans)
        out = (C.c_int32 * 4)()
        assert lib.kvc_launch_shape(_engine._SHAPE.pack(B, H, D, dtype, 0), n, rec, out) == 0
        return tuple(out)

    BF16, F32 = 2, 0
    c2 = shape_of(32, 32, 80, BF16, P.plan_fix_size([4096] * 32, 512, 0.2, "keep_low", [0, 1]))
    assert c2[:2] == (256, 3) and c2[3] == 0
    c3 = shape_of(32, 32, 80, BF16, P.plan_h2o([8192] * 32, 4, 64, 444, []))
    assert c3[:2] == (256, 3)
    c4 = shape_of(16, 8, 128, BF16, P.plan_snapkv([32768] * 32, 32, 512, 5, []))
    c5 = shape_of(8, 8, 128, BF16, P.plan_pyramid([32768] * 32, 512, 0.9, 64, "exponential", []))
    c5a = shape_of(8, 8, 128, BF16, P.plan_adaptive([32768] * 32, 512, 256, 1024, 0.3, 0.9, []))
    assert c4[:2] == c5[:2] == c5a[:2] == (512, 1) and c4[2] == 16
    c1 = shape_of(1, 32, 80, F32, P.plan_l2([2048] * 32, 0.8, 1000, [0, 1]))
    assert c1[:2] == (512, 1)                                   # few large units: short last wave
    steady_b1 = shape_of(1, 32, 80, BF16, P.plan_fix_size([513] * 32, 512, 0.2, "keep_low", [0, 1]))
    assert steady_b1[:2] == (256, 3)                            # small units: residency first
    big = shape_of(2, 8, 128, BF16, P.plan_h2o([200000], 4, 64, 444, []))
    assert big[3] == 1                                          # keys beyond shared memory: workspace
