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
    assert lib.kvc_abi_version() == 4
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
    io = _engine._IO.pack(16, 16, 16, 16, 120, 120, 12, 120, 120, 12, 0, 0, 0)
    assert lib.kvc_compress_layers(shape, 1, plan, io, None) == 2
    shape = _engine._SHAPE.pack(1, 1, 16, 2, 0)
    bad_plan = _engine._PLAN.pack(10, 1, 0, 20, 3, 1, 1, 1)  # sel_hi > seq_len
    assert lib.kvc_compress_layers(shape, 1, bad_plan, io, None) == 1


def test_structs_match_header_sizes():
    assert _engine._PLAN.size == 32 and _engine._IO.size == 104 and _engine._SHAPE.size == 20


def test_vote_workspace_sizing_is_opt_in_and_grows_with_the_sequence(monkeypatch):
    """kvc_vote_workspace_bytes is pure host arithmetic: 0 unless the split-sequence form is asked for, then one
    1 KB row-statistics record per (unit, slice) plus per-unit counters and final rows."""
    lib = _engine.load_library()

    def need(B, H, D, S, n_layers=2):
        shape = _engine._SHAPE.pack(B, H, D, 2, 0)
        layer = _engine._VOTE.pack(16, 16, 16, H * S * D, S * D, D, 4 * H * 32 * D, 32 * D, D, S, 0)
        return int(lib.kvc_vote_workspace_bytes(shape, n_layers, layer * n_layers))

    monkeypatch.delenv("KVC_VOTE_SPLIT", raising=False)
    assert need(2, 8, 128, 32768) == 0
    monkeypatch.setenv("KVC_VOTE_SPLIT", "1")
    monkeypatch.setenv("KVC_VOTE_TS", "8")
    small, big = need(2, 8, 128, 4096), need(2, 8, 128, 32768)
    units = 2 * 8 * 2
    assert small >= units * (4096 // 128 // 8) * 1024 and big >= units * (32768 // 128 // 8) * 1024
    assert small < big and big % 256 == 0
    assert need(2, 8, 96, 4096) == 0                       # head_dim 96: no vote kernel for that row width
    shape32 = _engine._SHAPE.pack(2, 8, 128, 0, 0)        # fp32 caches: the vote runs on 16-bit data only
    layer = _engine._VOTE.pack(16, 16, 16, 1, 1, 128, 1, 1, 128, 4096, 0)
    assert lib.kvc_vote_workspace_bytes(shape32, 1, layer) == 0
