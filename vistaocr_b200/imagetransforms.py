"""Device-side line-image pre-processing (reference src/imagetransforms.py).

Two surfaces over one kernel (csrc/preproc.cu, `vocr_scale_lines_u8`):

* the reference's transform classes with the same names, constructor arguments and call convention - `ConvertGray`,
  `Scale(new_h=...)`, `InvertBlackWhite`, `ToTensor`, `Compose` (imagetransforms.py:411-416,453-507,383-385,423-434,
  23-31) - so the pipelines of decode_testset.py:48-65 / train_cnn_lstm.py:263-279 / :301 read the same.  The classes
  record what they would do; `ToTensor`, the step that leaves uint8, runs the fused kernel on ONE image and returns the
  `[1,H,W]` float tensor (on the GPU).
* `LineBatchPreprocessor`: the batch form - B raw uint8 images in, the padded, (optionally width-sorted) float batch
  `[B,1,H,Wmax]` on the GPU out, in one upload of the raw bytes and one launch.  It also applies the 15-px width floor of
  OcrDataset.__getitem__ (ocr_dataset.py:174-180) and returns the widths CnnOcrModel.forward needs.

`Scale` reproduces what the reference's `cv2.resize(img, (w, h), self.interpolation)` really computes: OpenCV's default
8-bit INTER_LINEAR (the class's INTER_CUBIC lands on cv2.resize's positional `dst` parameter).  Only the `new_h=` form
with aspect-ratio preservation - the only one on the reference's line-recognition path - is implemented.
"""
import numpy as np
import torch

from . import _lib
from ._lib import check, lib, ptr, stream


def scaled_width(h, w, new_h):
    """imagetransforms.py:470-478: int(w * float(new_h / h)) in Python floats, non-positive -> 1."""
    nw = int(w * float(new_h / h))
    return nw if nw > 0 else 1


class _Pending:
    """A raw image travelling through the transform chain: the uint8 pixels plus the steps recorded so far."""

    def __init__(self, img):
        self.img = img
        self.gray = False
        self.new_h = None
        self.invert = False

    @property
    def shape(self):  # what downstream reference code may look at
        h, w = self.img.shape[:2]
        if self.new_h is not None:
            return (self.new_h, scaled_width(h, w, self.new_h))
        return (h, w)


def _pending(x):
    return x if isinstance(x, _Pending) else _Pending(np.ascontiguousarray(x))


class Compose:
    def __init__(self, transforms):
        self.transforms = transforms

    def __call__(self, img):
        for t in self.transforms:
            img = t(img)
        return img


class ConvertGray:
    def __call__(self, img):
        p = _pending(img)
        p.gray = True
        return p


class Scale:
    def __init__(self, size=None, new_h=None, new_w=None, preserve_apsect_ratio=True, interpolation=None):
        if size is not None or new_w is not None or new_h is None or not preserve_apsect_ratio:
            raise NotImplementedError("only Scale(new_h=H) with aspect-ratio preservation is on the line-recognition path")
        self.new_h = int(new_h)

    def __call__(self, img):
        p = _pending(img)
        if p.new_h is not None:
            raise NotImplementedError("two Scale steps in one chain")
        p.new_h = self.new_h
        return p


class InvertBlackWhite:
    def __call__(self, img):
        p = _pending(img)
        if p.new_h is None:
            raise NotImplementedError("InvertBlackWhite before Scale")
        p.invert = not p.invert
        return p


class ToTensor:
    """Runs the recorded chain on the device and returns float32 [1,H,W] in [0,1] (CUDA)."""

    def __init__(self, device="cuda"):
        self.device = torch.device(device)

    def __call__(self, pic):
        p = _pending(pic)
        if p.new_h is None:  # ToTensor alone: uint8 -> float / 255 through the same kernel (identity resize)
            p.new_h = p.img.shape[0]
        out, _ = _run_batch([p.img], p.new_h, p.invert, min_width=0, device=self.device, sort=False, gray=p.gray)
        return out[0]


def _run_batch(images, new_h, invert, min_width, device, sort, gray=True):
    B = len(images)
    if B == 0:
        raise ValueError("empty batch")
    chans = set()
    hs, ws, flat = np.zeros(B, np.int32), np.zeros(B, np.int32), []
    for i, im in enumerate(images):
        im = np.ascontiguousarray(im)
        if im.dtype != np.uint8:
            raise TypeError("raw uint8 images expected")
        if im.ndim == 3 and im.shape[2] == 4:  # BGRA: drop alpha (ocr_dataset.py:166-168)
            im = np.ascontiguousarray(im[:, :, :3])
        if im.ndim == 3 and im.shape[2] == 1:
            im = im[:, :, 0]
        if im.ndim == 3 and not gray:
            raise NotImplementedError("3-channel model input: add ConvertGray (the models on this path are 1-channel)")
        chans.add(1 if im.ndim == 2 else 3)
        if im.shape[0] == 0 or im.shape[1] == 0:
            raise ValueError("line image %d has zero area" % i)
        hs[i], ws[i] = im.shape[0], im.shape[1]
        flat.append(im.reshape(-1))
    if len(chans) != 1:
        raise ValueError("gray and colour images in one batch")
    channels = chans.pop()
    dws = np.array([scaled_width(int(h), int(w), new_h) for h, w in zip(hs, ws)], np.int32)
    widths = np.maximum(dws, min_width).astype(np.int32)
    order = np.argsort(-widths.astype(np.int64), kind="stable").astype(np.int32) if sort else np.arange(B, dtype=np.int32)
    offs = np.zeros(B, np.int64)
    offs[1:] = np.cumsum(hs[:-1].astype(np.int64) * ws[:-1] * channels)
    packed = torch.from_numpy(np.concatenate(flat)).pin_memory()
    w_out = int(widths.max())
    d_packed = packed.to(device, non_blocking=True)
    meta = torch.from_numpy(np.concatenate([hs, ws, dws, order])).to(device, non_blocking=True)
    d_offs = torch.from_numpy(offs).to(device, non_blocking=True)
    out = torch.empty((B, 1, new_h, w_out), dtype=torch.float32, device=device)
    st = lib().vocr_scale_lines_u8(ptr(d_packed), ptr(d_offs), ptr(meta[0:B]), ptr(meta[B:2 * B]), ptr(meta[2 * B:3 * B]),
                                   ptr(meta[3 * B:4 * B]), B, channels, new_h, w_out, int(bool(invert)), int(min_width),
                                   ptr(out), stream())
    check(st, "vocr_scale_lines_u8")
    return out, (widths, order)


class LineBatchPreprocessor:
    """images: list of raw uint8 arrays [h,w] (or [h,w,3] BGR with `gray=True` semantics of ConvertGray).
    Returns (batch float32 [B,1,H,Wmax] CUDA, widths int32[B] CPU tensor in batch order, order int32[B]) where
    order[b] = index of the image at batch position b (stable descending width when sort=True, like
    SortByWidthCollater, datautils.py:72)."""

    def __init__(self, line_height, invert=True, min_width=15, sort=True, device="cuda"):
        self.line_height = int(line_height)
        self.invert = bool(invert)
        self.min_width = int(min_width)
        self.sort = bool(sort)
        self.device = torch.device(device)

    def __call__(self, images):
        out, (widths, order) = _run_batch(images, self.line_height, self.invert, self.min_width, self.device, self.sort)
        return out, torch.from_numpy(widths[order].copy()), torch.from_numpy(order.copy())
