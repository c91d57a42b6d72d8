"""CPU: the C-ABI library builds, loads without a GPU and exports every symbol include/vistaocr_b200.h declares
(no compute calls here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "vistaocr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vocr_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from vistaocr_b200.build import build
    lib = ctypes.CDLL(build())
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "missing export: %s" % n


def test_ctypes_prototypes_cover_the_header():
    from vistaocr_b200 import _lib
    assert sorted(_lib.PROTOTYPES) == _declared()
    l = _lib.lib()
    assert l.vocr_version() >= 1000
    assert l.vocr_status_string(0) == b"success" and l.vocr_status_string(2) == b"invalid value"
    # sizing helpers are pure host code
    assert l.vocr_ctc_workspace_size(100, 4, 80, 20) > 0
    assert l.vocr_bilstm_workspace_size(64, 512, 0) > 0 and l.vocr_bilstm_workspace_size(64, 513, 0) == 0


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "vistaocr_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py") and fn != "_smoke.py":
            assert "oracle" not in open(os.path.join(pkg, fn)).read(), fn
