"""CPU checks of the C-ABI boundary: the library loads and exports every symbol include/msx.h declares
(no compute calls without a GPU), and argument validation fails loudly through msx_last_error()."""
import ctypes
import os
import re

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libmsx():
    import __graft_entry__ as ge
    if not os.path.exists(ge.LIB):
        ge.build()
    from musicstyletransfer_b200 import lib
    return lib.load()


def declared_symbols():
    src = open(os.path.join(REPO, "include", "msx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(msx_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(libmsx):
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(libmsx, n), "include/msx.h declares %s but libmsx.so does not export it" % n


def test_python_binding_uses_declared_symbols():
    """Every symbol ops.py / featurise.py binds is declared in the header."""
    declared = set(declared_symbols())
    used = set()
    for f in ("ops.py", "featurise.py", "sampling.py"):
        path = os.path.join(REPO, "musicstyletransfer_b200", f)
        if os.path.exists(path):
            used |= set(re.findall(r'"(msx_[a-z0-9_]+)"', open(path).read()))
    assert used and used <= declared, used - declared


def test_version_and_error_reporting(libmsx):
    assert libmsx.msx_version() >= 100
    libmsx.msx_last_error.restype = ctypes.c_char_p
    # null pointers are rejected before any CUDA call
    rc = libmsx.msx_rasterize(None, None, None, None, 4, 120, 4, 64, 64, 0, None, None, None, None)
    assert rc == -1
    assert b"null pointer" in libmsx.msx_last_error()
    rc = libmsx.msx_gemm_f32(None, 0, 0, None, 0, 0, None, 0, 4, 4, 4, None, 0, ctypes.c_float(0.0),
                             ctypes.c_ulonglong(0), 0, None, 0, ctypes.c_float(1.0), 0, 1, None, None)
    assert rc == -1


def test_product_does_not_import_oracle():
    """The product package must never route through the oracle (only smoke.py, the checker, may)."""
    pkg = os.path.join(REPO, "musicstyletransfer_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py") and f != "smoke.py":
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(root, f)


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    from musicstyletransfer_b200 import lib
    monkeypatch.setattr(lib, "_lib", None)
    monkeypatch.setattr(lib, "LIB_PATH", str(tmp_path / "libmsx.so"))
    with pytest.raises(lib.MsxError):
        lib.load()
