"""Device-side `SortByWidthCollater` (reference src/datautils.py:61-176).

Same call and return convention as the reference collater - `collater(batch)` with `batch` a list of
`(image[C,H,w] float tensor, transcript list[int], metadata dict with 'width')` returns
`(input_tensor[B,C,H,Wmax], target int32 1-D, input_widths int32[B], target_widths int32[B], metadata)` with samples
sorted by `metadata['width']` descending (stable) - but the padded batch is assembled ON THE GPU: only the ragged
pixels cross PCIe, the zero padding is written by the copy kernel.  `input_tensor` comes back as a CUDA tensor (the
reference's `train()` then calls `.cuda()` on it, a no-op); the small integer tensors stay on the CPU exactly like the
reference's, because CnnOcrModel.forward and the CTC wrapper read them there.
"""
import numpy as np
import torch

from . import _lib
from ._lib import check, lib, ptr, stream


def stable_desc_order(keys):
    """Order of `list.sort(key=..., reverse=True)`: descending, equal keys keep their original order."""
    keys = np.asarray(keys)
    return np.argsort(-keys.astype(np.int64), kind="stable").astype(np.int32)


class SortByWidthCollater:
    def __init__(self, device="cuda"):
        self.device = torch.device(device)

    def __call__(self, batch):
        B = len(batch)
        if B == 0:
            raise ValueError("empty batch")
        keys = [int(m["width"]) for _, _, m in batch]
        order = stable_desc_order(keys)
        c, h = int(batch[0][0].size(0)), int(batch[0][0].size(1))
        img_w = np.array([int(t.size(2)) for t, _, _ in batch], np.int32)
        w_out = int(img_w[order[0]])  # the reference sizes the batch by the first tensor after sorting
        if (img_w > w_out).any():
            raise RuntimeError("an image is wider than the widest-by-key image (the reference fails here too)")
        offs = np.zeros(B, np.int64)
        offs[1:] = np.cumsum(img_w[:-1].astype(np.int64) * c * h)
        packed = torch.cat([t.reshape(-1).float() for t, _, _ in batch]).pin_memory()
        lab_lens = np.array([len(tr) for _, tr, _ in batch], np.int32)
        lab_offs = np.zeros(B + 1, np.int32)
        lab_offs[1:] = np.cumsum(lab_lens)
        flat_labels = np.fromiter((ch for _, tr, _ in batch for ch in tr), dtype=np.int32, count=int(lab_offs[-1]))
        dev = self.device
        d_packed = packed.to(dev, non_blocking=True)
        d_offs = torch.from_numpy(offs).to(dev, non_blocking=True)
        d_w = torch.from_numpy(img_w).to(dev, non_blocking=True)
        d_order = torch.from_numpy(order).to(dev, non_blocking=True)
        d_lab = torch.from_numpy(flat_labels if flat_labels.size else np.zeros(1, np.int32)).to(dev, non_blocking=True)
        d_lab_offs = torch.from_numpy(lab_offs).to(dev, non_blocking=True)
        out = torch.empty((B, c, h, w_out), dtype=torch.float32, device=dev)
        d_labels_out = torch.empty((max(1, int(lab_offs[-1])),), dtype=torch.int32, device=dev)
        d_lens_out = torch.empty((B,), dtype=torch.int32, device=dev)
        st = lib().vocr_collate_lines_f32(ptr(d_packed), ptr(d_offs), ptr(d_w), ptr(d_order), B, c, h, w_out, ptr(out),
                                          ptr(d_lab), ptr(d_lab_offs), ptr(d_labels_out), ptr(d_lens_out), stream())
        check(st, "vocr_collate_lines_f32")
        # host copies of the small integer outputs (what the reference returns on the CPU)
        input_widths = torch.from_numpy(np.asarray(keys, np.int32)[order].copy())
        target_widths = torch.from_numpy(lab_lens[order].copy())
        target = torch.from_numpy(np.concatenate([np.asarray(batch[i][1], np.int32).reshape(-1) for i in order])
                                  if lab_offs[-1] > 0 else np.zeros(0, np.int32))
        metadata = {}
        if any("writer-id" in m for _, _, m in batch):
            metadata["writer-ids"] = torch.tensor([batch[i][2].get("writer-id", 0) for i in order], dtype=torch.long)
        if any("utt-id" in m for _, _, m in batch):
            metadata["utt-ids"] = [batch[i][2]["utt-id"] for i in order if "utt-id" in batch[i][2]]
        metadata["device_target"] = d_labels_out[:int(lab_offs[-1])]
        metadata["device_target_widths"] = d_lens_out
        return out, target, input_widths, target_widths, metadata
