"""No-GPU checks of the C-ABI boundary: the library loads, exports every symbol include/*.h
declares, and fails loudly (no fallback) when there is no device."""
import ctypes
import re
from pathlib import Path

import pytest

from xfeatslam_b200 import capi

REPO = Path(__file__).resolve().parents[1]


def declared_symbols():
    text = (REPO / "include" / "xfeat_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(xfb_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(capi.SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = capi.load_library()
    for s in declared_symbols():
        assert hasattr(lib, s), s


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.XFBError) as ei:
        capi.XFeatB200(max_h=64, max_w=64)
    assert "no usable CUDA device" in str(ei.value) or "CUDA" in str(ei.value)


def test_create_rejects_bad_arguments():
    lib = capi.load_library()
    h = ctypes.c_void_p()
    assert lib.xfb_create(ctypes.byref(h), None, 0, 0, 480, 640, 1, 4096) == -1      # XFB_ERR_ARG
    blob = b"XFBW" + b"\0" * 60
    assert lib.xfb_create(ctypes.byref(h), blob, len(blob), 0, 16, 16, 1, 4096) == -1  # image too small
    assert lib.xfb_create(ctypes.byref(h), blob, len(blob), 0, 480, 640, 1, 100000) == -1
    assert lib.xfb_last_error(None) is not None


def test_product_never_imports_oracle():
    """The shipped package must not import / include / dlopen anything under oracle/
    (only tests, smoke and bench may); comments that *cite* the oracle are fine."""
    pat = re.compile(r"(^\s*(from|import)\s+oracle\b)|(#\s*include\s*[\"<][^\">]*oracle)|(libmatcher_oracle)|(oracle[/.]_ref)|(ref_xfeat)", re.M)
    for p in (REPO / "xfeatslam_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".h", ".cc", ".cuh", ".cpp"):
            assert not pat.search(p.read_text()), p
