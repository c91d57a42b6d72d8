"""Oracle (test infrastructure): SortByWidthCollater restated in numpy (reference src/datautils.py:61-176,
non-seq2seq, non-bylang branch).  Pinned against the reference function itself in tests/test_oracle_vs_reference.py."""
import numpy as np


def collate_ref(batch):
    """batch: list of (image ndarray [C,H,w], transcript list[int], metadata dict).  Returns
    (input [B,C,H,W0] float32, target int32, input_widths int32, target_widths int32, order)."""
    order = sorted(range(len(batch)), key=lambda i: batch[i][2]["width"], reverse=True)  # python sort is stable
    first = batch[order[0]][0]
    out = np.zeros((len(batch),) + first.shape, np.float32)
    widths = np.zeros(len(batch), np.int32)
    tw = np.zeros(len(batch), np.int32)
    target = []
    for idx, i in enumerate(order):
        img, tr, md = batch[i]
        out[idx, :, :, :img.shape[2]] = img
        widths[idx] = md["width"]
        tw[idx] = len(tr)
        target.extend(tr)
    return out, np.array(target, np.int32), widths, tw, np.array(order, np.int32)
