"""Generates tests/golden/preproc.npz: raw uint8 line images and what the reference's pre-processing chain
Scale(new_h) -> InvertBlackWhite -> ToTensor produces for them, computed with cv2 exactly as the reference calls it
(src/imagetransforms.py:478: cv2.resize(img, (w, h), self.interpolation) - interpolation passed positionally).
Run in the authoring container (needs cv2):  python tests/golden/make_golden_preproc.py"""
import os

import cv2
import numpy as np

rng = np.random.default_rng(7)
cases = [(47, 301, 30), (60, 400, 30), (33, 20, 30), (120, 733, 60), (30, 211, 30), (25, 90, 60), (61, 9, 30)]
out = {}
for k, (h, w, new_h) in enumerate(cases):
    img = rng.integers(0, 256, size=(h, w), dtype=np.uint8)
    nw = int(w * float(new_h / h)) or 1
    r = cv2.resize(img, (nw, new_h), cv2.INTER_CUBIC)           # the reference's call
    t = (-r + 255).astype(np.float32) / np.float32(255)         # InvertBlackWhite, ToTensor
    out["img%d" % k] = img
    out["ref%d" % k] = t
    out["new_h%d" % k] = np.int32(new_h)
bgr = rng.integers(0, 256, size=(40, 123, 3), dtype=np.uint8)
out["bgr"] = bgr
out["bgr_gray"] = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "preproc.npz"), **out)
print("wrote", len(cases), "cases")
