"""ORACLE (test infrastructure only - never imported by the product path).

CPU restatement of the image pre-processing in front of the hot path:
    imagetransforms.Scale(new_h=H)   reference src/imagetransforms.py:453-507
    imagetransforms.InvertBlackWhite reference src/imagetransforms.py:383-385      (-img + 255 on uint8)
    imagetransforms.ToTensor         reference src/imagetransforms.py:423-434      (.float().div(255))
    width floor of 15 px, padded with ones: reference src/ocr_dataset.py:174-180

What `Scale` really computes.  The reference calls `cv2.resize(img, (w, h), self.interpolation)` (:478, :489-498):
the third POSITIONAL parameter of cv2.resize is `dst`, not `interpolation`, so the INTER_CUBIC default of the class
never reaches OpenCV and every resize runs with OpenCV's default, INTER_LINEAR (checked against cv2 in
tests/test_oracle_preproc.py: positional == INTER_LINEAR, != INTER_CUBIC).  The algorithm lives in a third-party
dependency that is not under /root/reference: OpenCV (cv2, unpinned by the reference; 4.13.0 in this image),
modules/imgproc/src/resize.cpp.  Its 8-bit bilinear path is pure integer arithmetic, restated here:
  * scale = 1 / (dst / src) in float64;  fx = float32((dx + 0.5) * scale - 0.5);  sx = floor(fx);  fx -= sx
    left border: sx < 0 -> (sx, fx) = (0, 0);  right border: sx >= w - 1 -> (sx, fx) = (w - 1, 0)   (x only; in y the
    two source rows are clamped to [0, h - 1] and the weights are kept)
  * weights are 11-bit fixed point: a1 = round_half_even(fx * 2048), a0 = round_half_even((1 - fx) * 2048) (int16)
  * horizontal pass (int32): D = S[sx] * a0 + S[sx + 1] * a1
  * vertical pass: out = ((b0 * (D0 >> 4)) >> 16) + ((b1 * (D1 >> 4)) >> 16) + 2) >> 2
  * exact 2x down-scaling in both directions (src = 2 * dst) is re-routed to INTER_AREA: (s00 + s01 + s10 + s11 + 2) >> 2
  * dst size == src size: copy.
The pin is cv2 itself (same image on the GPU box) plus the committed fixture tests/golden/preproc.npz.
"""
import numpy as np


def scaled_width(h, w, new_h):
    """Scale(new_h=H) with preserve_aspect_ratio: local_new_w = int(w * float(new_h / h)) (imagetransforms.py:470-473);
    a non-positive result falls back to 1 (:475-478)."""
    nw = int(w * float(new_h / h))
    return nw if nw > 0 else 1


def _coeffs(dst, src):
    """Per output index: first source index and the two int16 weights of OpenCV's 8-bit INTER_LINEAR."""
    inv_scale = float(dst) / float(src)
    scale = 1.0 / inv_scale
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int32)
    f = (f - s.astype(np.float32)).astype(np.float32)
    return s, f


def _fix(f):
    a1 = np.rint(f * np.float32(2048)).astype(np.int32)          # cvRound: round half to even
    a0 = np.rint((np.float32(1.0) - f) * np.float32(2048)).astype(np.int32)
    return np.clip(a0, -32768, 32767), np.clip(a1, -32768, 32767)


def resize_linear_u8(img, dst_w, dst_h):
    """cv2.resize(img, (dst_w, dst_h)) for a 2-D uint8 image (default interpolation = INTER_LINEAR)."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape
    if (dst_w, dst_h) == (w, h):
        return img.copy()
    sx_scale = 1.0 / (float(dst_w) / float(w))
    sy_scale = 1.0 / (float(dst_h) / float(h))
    if abs(sx_scale - 2) < np.finfo(np.float64).eps and abs(sy_scale - 2) < np.finfo(np.float64).eps:
        s = img.astype(np.int32)
        return ((s[0:2 * dst_h:2, 0:2 * dst_w:2] + s[0:2 * dst_h:2, 1:2 * dst_w:2] + s[1:2 * dst_h:2, 0:2 * dst_w:2] +
                 s[1:2 * dst_h:2, 1:2 * dst_w:2] + 2) >> 2).astype(np.uint8)
    sx, fx = _coeffs(dst_w, w)
    lo = sx < 0
    sx[lo], fx[lo] = 0, 0
    hi = sx >= w - 1
    sx[hi], fx[hi] = w - 1, 0
    a0, a1 = _fix(fx)
    sy, fy = _coeffs(dst_h, h)
    b0, b1 = _fix(fy)
    r0 = np.clip(sy, 0, h - 1)
    r1 = np.clip(sy + 1, 0, h - 1)
    s = img.astype(np.int32)
    sx1 = np.minimum(sx + 1, w - 1)  # weight 0 wherever this clamps
    hpass = s[:, sx] * a0[None, :] + s[:, sx1] * a1[None, :]     # [h, dst_w] int32
    d0, d1 = hpass[r0], hpass[r1]
    out = (((b0[:, None] * (d0 >> 4)) >> 16) + ((b1[:, None] * (d1 >> 4)) >> 16) + 2) >> 2
    return out.astype(np.uint8)


def preprocess_line(img, new_h, invert=True, min_width=15):
    """Scale(new_h) -> [InvertBlackWhite] -> ToTensor -> width floor (ones), as float32 [1, new_h, max(w', min_width)]."""
    h, w = img.shape
    nw = scaled_width(h, w, new_h)
    r = resize_linear_u8(img, nw, new_h)
    if invert:
        r = (255 - r.astype(np.int32)).astype(np.uint8)          # -img + 255 in uint8 arithmetic
    t = r.astype(np.float32) / np.float32(255)
    if nw < min_width:
        t2 = np.ones((new_h, min_width), np.float32)
        t2[:, :nw] = t
        t = t2
    return t[None]
