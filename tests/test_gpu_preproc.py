"""GPU parity: csrc/preproc.cu through the C ABI (vistaocr_b200/imagetransforms.py) against oracle/preproc_ref.py,
itself pinned to cv2.  Bit-exact: the resize is integer work, the final division by 255 is one IEEE operation."""
import os

import numpy as np
import pytest
import torch

from oracle.preproc_ref import preprocess_line, scaled_width

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "preproc.npz")


def _images(rng, B, new_h):
    imgs = []
    for i in range(B):
        h, w = int(rng.integers(max(4, new_h // 3), 3 * new_h)), int(rng.integers(3, 900))
        if i % 5 == 0:
            h, w = 2 * new_h, 2 * int(rng.integers(8, 400))  # OpenCV's exact-2x area path
        if i % 7 == 1:
            h = new_h                                           # no vertical change
        imgs.append(rng.integers(0, 256, size=(h, w), dtype=np.uint8))
    return imgs


@pytest.mark.parametrize("new_h,B,invert", [(30, 17, True), (60, 9, True), (120, 6, False)])
def test_batch_matches_oracle(cuda, new_h, B, invert):
    from vistaocr_b200.imagetransforms import LineBatchPreprocessor
    rng = np.random.default_rng(new_h + B)
    imgs = _images(rng, B, new_h)
    imgs.append(rng.integers(0, 256, size=(new_h * 2, 12), dtype=np.uint8))  # narrower than the 15-px floor
    out, widths, order = LineBatchPreprocessor(new_h, invert=invert, device=cuda)(imgs)
    out = out.cpu().numpy()
    want_w = np.array([max(15, scaled_width(im.shape[0], im.shape[1], new_h)) for im in imgs])
    assert np.array_equal(order.numpy(), np.argsort(-want_w, kind="stable"))  # stable descending, datautils.py:72
    assert np.array_equal(widths.numpy(), want_w[order.numpy()])
    assert out.shape == (len(imgs), 1, new_h, want_w.max())
    for b, i in enumerate(order.numpy()):
        ref = preprocess_line(imgs[i], new_h, invert=invert, min_width=15)
        w = ref.shape[2]
        assert np.array_equal(out[b, :, :, :w], ref), (b, i)
        assert not out[b, :, :, w:].any()  # batch padding is exact zeros


def test_golden_and_transform_classes(cuda):
    from vistaocr_b200 import imagetransforms as it
    g = np.load(GOLD)
    k = 0
    while "img%d" % k in g:
        new_h = int(g["new_h%d" % k])
        chain = it.Compose([it.Scale(new_h=new_h), it.InvertBlackWhite(), it.ToTensor(cuda)])  # decode_testset.py:48-65
        t = chain(g["img%d" % k])
        assert t.is_cuda and t.dtype == torch.float32 and tuple(t.shape) == (1,) + g["ref%d" % k].shape
        assert np.array_equal(t.cpu().numpy()[0], g["ref%d" % k]), k
        k += 1
    # validation-set chain of train_cnn_lstm.py:301 (no inversion) and ConvertGray on a BGR image
    t = it.Compose([it.Scale(new_h=30), it.ToTensor(cuda)])(g["img0"])
    assert np.array_equal(t.cpu().numpy(), preprocess_line(g["img0"], 30, invert=False, min_width=0))
    t = it.Compose([it.ConvertGray(), it.Scale(new_h=30), it.InvertBlackWhite(), it.ToTensor(cuda)])(g["bgr"])
    assert np.array_equal(t.cpu().numpy(), preprocess_line(g["bgr_gray"], 30, invert=True, min_width=0))
    t = it.ToTensor(cuda)(g["img1"])  # ToTensor alone: v / 255
    assert np.array_equal(t.cpu().numpy()[0], g["img1"].astype(np.float32) / np.float32(255))


def test_full_size_properties(cuda):
    """cfg3-sized batch (height 120 -> widths up to 2000): properties that need no CPU pass over every pixel."""
    from vistaocr_b200.imagetransforms import LineBatchPreprocessor
    rng = np.random.default_rng(5)
    imgs = [rng.integers(0, 256, size=(120, int(w)), dtype=np.uint8) for w in rng.integers(400, 2001, size=64)]
    pre = LineBatchPreprocessor(120, invert=True, sort=False, device=cuda)
    out, widths, _ = pre(imgs)            # same size in and out: the chain reduces to (255 - v) / 255
    for b in (0, 17, 63):
        assert np.array_equal(out[b, 0, :, :imgs[b].shape[1]].cpu().numpy(), (255 - imgs[b].astype(np.float32)) / np.float32(255))
    plain, _, _ = LineBatchPreprocessor(120, invert=False, sort=False, device=cuda)(imgs)
    valid = torch.arange(out.shape[3], device=cuda)[None, None, None, :] < widths.to(cuda)[:, None, None, None]
    assert torch.equal(torch.where(valid, (out * 255).round() + (plain * 255).round(), torch.full_like(out, 255.)),
                       torch.full_like(out, 255.))       # inversion is an involution on the byte values
    const = [np.full((97, 640), 200, np.uint8), np.full((240, 1000), 13, np.uint8)]
    o, w, _ = LineBatchPreprocessor(60, invert=False, sort=False, device=cuda)(const)
    assert torch.equal(o[0, 0, :, :int(w[0])], torch.full((60, int(w[0])), 200 / 255, device=cuda))  # constants survive
    assert torch.equal(o[1, 0, :, :int(w[1])], torch.full((60, int(w[1])), np.float32(13) / np.float32(255), device=cuda))
