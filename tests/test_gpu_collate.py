"""GPU parity (bit-exact) of the device-side collater against the oracle restatement of SortByWidthCollater."""
import numpy as np
import pytest
import torch

from oracle.collate_ref import collate_ref

pytestmark = pytest.mark.gpu


def _batch(rng, B, C, H, wmin, wmax, dup=True):
    out = []
    for i in range(B):
        w = int(rng.integers(wmin, wmax + 1))
        if dup and i % 5 == 4:
            w = out[-1][2]["width"]  # equal keys: stability matters
        img = rng.random((C, H, w), dtype=np.float32)
        tr = rng.integers(1, 90, size=int(rng.integers(0, 12))).tolist()
        out.append((img, tr, {"width": w, "utt-id": "utt%03d" % i, "writer-id": i}))
    return out


@pytest.mark.parametrize("B,C,H,wmin,wmax", [(64, 1, 30, 15, 300), (7, 3, 12, 1, 40), (1, 1, 60, 100, 100),
                                             (33, 1, 120, 40, 500)])
def test_collate_matches_oracle(cuda, B, C, H, wmin, wmax):
    from vistaocr_b200.datautils import SortByWidthCollater
    rng = np.random.default_rng(B * 7 + H)
    batch = _batch(rng, B, C, H, wmin, wmax)
    want = collate_ref(batch)
    got = SortByWidthCollater()([(torch.from_numpy(i), t, m) for i, t, m in batch])
    assert got[0].is_cuda and torch.equal(got[0].cpu(), torch.from_numpy(want[0]))
    assert got[1].dtype == torch.int32 and got[1].tolist() == want[1].tolist()
    assert got[2].tolist() == want[2].tolist() and got[3].tolist() == want[3].tolist()
    assert got[4]["device_target"].cpu().tolist() == want[1].tolist()
    assert got[4]["device_target_widths"].cpu().tolist() == want[3].tolist()
    assert got[4]["utt-ids"] == [batch[i][2]["utt-id"] for i in want[4]]


def test_padded_images_keep_their_tensor_width(cuda):
    """metadata['width'] may be smaller than the tensor (the dataset pads to 15 px, ocr_dataset.py:179-182)."""
    from vistaocr_b200.datautils import SortByWidthCollater
    rng = np.random.default_rng(0)
    batch = [(rng.random((1, 8, 15), dtype=np.float32), [1, 2], {"width": 9}),
             (rng.random((1, 8, 15), dtype=np.float32), [3], {"width": 12})]
    want = collate_ref(batch)
    got = SortByWidthCollater()([(torch.from_numpy(i), t, m) for i, t, m in batch])
    assert torch.equal(got[0].cpu(), torch.from_numpy(want[0])) and got[2].tolist() == [12, 9]
