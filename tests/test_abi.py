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
    assert l.vocr_bilstm_workspace_size(100, 64, 512, 0) > 0 and l.vocr_bilstm_workspace_size(100, 64, 513, 0) == 0


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "vistaocr_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py") and fn != "_smoke.py":
            assert "oracle" not in open(os.path.join(pkg, fn)).read(), fn


def test_precision_switch_is_host_only_state():
    import pytest
    import vistaocr_b200
    assert vistaocr_b200.get_precision() == "fp32"
    assert vistaocr_b200.set_precision("fp16") == "fp32" and vistaocr_b200.get_precision() == "fp16"
    assert vistaocr_b200.set_precision("fp32") == "fp16" and vistaocr_b200.get_precision() == "fp32"
    with pytest.raises(ValueError):
        vistaocr_b200.set_precision("bf16")
    from vistaocr_b200 import _lib
    assert _lib.lib().vocr_set_tc_products(2) != 0 and _lib.lib().vocr_get_tc_products() == 3


def test_scaled_width_matches_the_reference_formula():
    """imagetransforms.py:470-478: int(w * float(new_h / h)) in Python floats; the product's host helper and the
    oracle's must agree everywhere (the kernel takes the width from the host)."""
    from oracle.preproc_ref import scaled_width as ref
    from vistaocr_b200.imagetransforms import scaled_width
    for h in (7, 24, 30, 47, 60, 61, 120, 333):
        for w in (1, 2, 9, 15, 100, 301, 733, 1999, 2000):
            for new_h in (30, 60, 120):
                assert scaled_width(h, w, new_h) == ref(h, w, new_h) == max(1, int(w * float(new_h / h)))
