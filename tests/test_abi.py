"""CPU: the C-ABI library loads and exports every symbol include/ovqa.h declares; the Python binding declares
every one of them; failures are loud (no CPU fallback).  No compute calls here."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "ovqa.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ovqa_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built_lib():
    import __graft_entry__ as g
    g.build()
    from openvivqa_b200 import _lib
    return _lib


def test_header_declares_the_hot_path_entry_points():
    syms = header_symbols()
    for required in ("ovqa_gemm", "ovqa_sdpa_fwd", "ovqa_sdpa_bwd", "ovqa_layernorm_fwd", "ovqa_layernorm_bwd",
                     "ovqa_ptr_scores", "ovqa_prev_pred_embed", "ovqa_greedy_step", "ovqa_box_geometry_bias"):
        assert required in syms


def test_library_exports_every_declared_symbol(built_lib):
    raw = ctypes.CDLL(built_lib.LIB_PATH)
    for s in header_symbols():
        assert hasattr(raw, s), f"{s} declared in include/ovqa.h but not exported"


def test_binding_covers_the_header(built_lib):
    assert sorted(built_lib.declared_symbols()) == header_symbols()


def test_version_and_error_strings(built_lib):
    l = built_lib.lib()
    assert l.ovqa_version() >= 100
    assert l.ovqa_error_string(0) == b"ok"
    for code in (-1, -2, -3, -4, -5):
        assert len(l.ovqa_error_string(code)) > 3


def test_bad_arguments_return_codes_not_crashes(built_lib):
    l = built_lib.lib()
    # null pointers / bad dtype are rejected before any CUDA call
    assert l.ovqa_gemm(None, None, None, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 1.0, 0.0, None, 0, None, None, 0, None, 0, None, 0,
                       0.0, 0, 0, None, None, 1, None) == -1
    assert l.ovqa_layernorm_fwd(None, None, None, None, None, 0, None, None, None, None, None, 4, 512, 1e-5, None) == -1
    with pytest.raises(built_lib.OvqaError):
        built_lib.check(-2, "x")


def test_cpu_tensors_are_an_error_not_a_fallback(built_lib):
    from openvivqa_b200 import ops
    a = torch.randn(4, 8)
    with pytest.raises(ValueError):
        ops.gemm(a, a)
    with pytest.raises(ValueError):
        ops.layernorm_fwd(a, torch.ones(8), torch.zeros(8))


def test_missing_library_fails_loudly(monkeypatch, built_lib):
    monkeypatch.setattr(built_lib, "_lib", None)
    monkeypatch.setattr(built_lib, "LIB_PATH", "/nonexistent/libovqa_b200.so")
    with pytest.raises(built_lib.OvqaError):
        built_lib.lib()


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "openvivqa_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(d, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports oracle/"
