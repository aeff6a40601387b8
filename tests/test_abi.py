"""CPU checks of the drop-in boundary: the shared library loads, exports every symbol that
include/gdr.h declares, the ctypes table covers them, and the product never touches oracle/."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "graph-distillation-for-recommendation_b200")


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "gdr.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gdr_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    lib = ctypes.CDLL(os.path.join(PKG, "lib", "libgdr_b200.so"))
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/gdr.h but not exported"
    assert lib.gdr_abi_version() == 1


def test_ctypes_table_matches_header():
    from gdr import _lib
    assert sorted(_lib._SIGNATURES) == declared_symbols()


def test_argument_validation_without_gpu():
    """Entry points reject bad arguments before touching the device."""
    from gdr import _lib
    import pytest
    with pytest.raises(_lib.GdrError) as e:
        _lib.call("gdr_spmm_prop", 4, 6, 0, 0, 0, 1.0, 0, 8, 0, 8, 0, 0, 0.0, 0)
    assert e.value.code == -1 and "null" in str(e.value)
    with pytest.raises(_lib.GdrError):
        _lib.call("gdr_kmeans_assign", 10, 0, 4, 0, 4, 0, 4, 0, 0, 0, 0, 0, 0, 0, 0)
    assert _lib.query("gdr_sort_pairs_ws_bytes", 1000) > 12000


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+\.*oracle|import_module\(.oracle|liboracle", src, re.M), \
                    f"{f} reaches into oracle/"


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    from gdr import _lib
    import pytest
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(ImportError):
        _lib.load()
