"""CPU tests of the drop-in boundary: the C-ABI library loads here (no GPU), exports every symbol that
include/havc_b200.h declares, the ctypes mirror covers exactly those symbols, and argument validation fails
loudly without touching a device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "havc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(havc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from vsdeoldify_b200 import _lib
    lib = _lib.lib()
    names = _declared()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in havc_b200.h but not exported by libhavc_b200.so"
    assert sorted(_lib.exported_symbols()) == names, "ctypes signatures and header drifted apart"


def test_struct_layout_matches_header():
    from vsdeoldify_b200 import _lib
    # natural alignment on x86-64: havc_act_view = ptr + 5*int32 (+4 pad) + 4*int64 = 64 bytes
    assert ctypes.sizeof(_lib.ActView) == 64
    assert _lib.ConvDesc.src1.offset == 64 and _lib.ConvDesc.weight.offset == 128


def test_argument_validation_without_device():
    from vsdeoldify_b200 import _lib
    lib = _lib.lib()
    d = _lib.ConvDesc()
    assert lib.havc_conv_gemm(ctypes.byref(d), None) == -1           # HAVC_ERR_ARG, no kernel launched
    assert b"havc_conv_gemm" in lib.havc_last_error()
    assert lib.havc_blur2x2(None, None, 1, 8, 8, 8, 8, 0, None) == -1
    assert lib.havc_softmax_rows(None, 2, None, 4, 6, 6, 6, 0, None) == -1
    assert lib.havc_version() >= 100


def test_missing_library_fails_loudly(monkeypatch):
    from vsdeoldify_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", os.path.join(ROOT, "does_not_exist.so"))
    with pytest.raises(_lib.HavcLibraryError):
        _lib.lib()


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under vsdeoldify_b200/ may import it."""
    pkg = os.path.join(ROOT, "vsdeoldify_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports oracle/"
