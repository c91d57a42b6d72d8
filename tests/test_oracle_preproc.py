"""CPU: the pre-processing restatement (oracle/preproc_ref.py) against OpenCV itself - called the way the reference
calls it - and against the committed fixture tests/golden/preproc.npz.  Bit-exact (uint8 / identical float32)."""
import os

import numpy as np
import pytest

from oracle.preproc_ref import preprocess_line, resize_linear_u8, scaled_width

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "preproc.npz")


def test_golden_fixture():
    g = np.load(GOLD)
    k = 0
    while "img%d" % k in g:
        got = preprocess_line(g["img%d" % k], int(g["new_h%d" % k]), invert=True, min_width=0)
        assert got.dtype == np.float32 and np.array_equal(got[0], g["ref%d" % k]), k
        k += 1
    assert k >= 5
    b = g["bgr"].astype(np.int64)
    gray = ((b[..., 0] * 3735 + b[..., 1] * 19235 + b[..., 2] * 9798 + (1 << 14)) >> 15).astype(np.uint8)
    assert np.array_equal(gray, g["bgr_gray"])


def test_width_floor_pads_with_ones():
    img = np.full((60, 10), 255, np.uint8)  # white background -> 0 after inversion
    t = preprocess_line(img, 30, invert=True, min_width=15)
    assert t.shape == (1, 30, 15) and not t[0, :, :5].any() and (t[0, :, 5:] == 1).all()


def test_matches_cv2_as_the_reference_calls_it():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    for it in range(150):
        h, w = int(rng.integers(8, 160)), int(rng.integers(2, 700))
        new_h = int(rng.choice([30, 60, 120, 16, h]))
        if it % 6 == 0:  # exact 2x down-scaling: OpenCV's area path
            h, w = 2 * new_h, 2 * int(rng.integers(5, 300))
        img = rng.integers(0, 256, size=(h, w), dtype=np.uint8)
        nw = scaled_width(h, w, new_h)
        ref = cv2.resize(img, (nw, new_h), cv2.INTER_CUBIC)  # src/imagetransforms.py:478 (third positional = dst)
        assert np.array_equal(resize_linear_u8(img, nw, new_h), ref), (h, w, new_h)
    img = rng.integers(0, 256, size=(47, 301), dtype=np.uint8)
    a = cv2.resize(img, (192, 30), cv2.INTER_CUBIC)
    assert np.array_equal(a, cv2.resize(img, (192, 30), interpolation=cv2.INTER_LINEAR))
    assert not np.array_equal(a, cv2.resize(img, (192, 30), interpolation=cv2.INTER_CUBIC))
