"""Cycle counters of ctc_mitm_kernel (library built with VOCR_NVCC_FLAGS=-DVOCR_CTC_PROF): per warp of the first four
utterances, cycles spent before the meeting point, at the barrier and after it."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vistaocr_b200 import _lib  # noqa: E402


def a256(x):
    return (x + 255) & ~255


def main():
    T, B, A, L = [int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (500, 256, 120, 50))]
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(4)
    x = torch.randn((T, B, A), device=dev)
    act_lens = torch.from_numpy(np.sort(rng.integers(T // 2, T + 1, size=B))[::-1].astype(np.int32).copy()).to(dev)
    label_lens = torch.full((B,), L, dtype=torch.int32, device=dev)
    labels = torch.from_numpy(rng.integers(1, A, size=B * L).astype(np.int32)).to(dev)
    l = _lib.lib()
    nbytes = l.vocr_ctc_workspace_size(T, B, A, L)
    ws = torch.zeros((nbytes,), dtype=torch.uint8, device=dev)
    costs = torch.empty((B,), dtype=torch.float32, device=dev)
    grads = torch.empty_like(x)
    for _ in range(3):
        st = l.vocr_ctc_loss_f32(_lib.ptr(x), _lib.ptr(grads), _lib.ptr(labels), _lib.ptr(label_lens), _lib.ptr(act_lens),
                                 T, B, A, L, _lib.ptr(costs), _lib.ptr(ws), nbytes, _lib.stream())
        _lib.check(st, "ctc")
    torch.cuda.synchronize()
    base = (ws.data_ptr() + 255) & ~255
    off = base - ws.data_ptr()
    off += a256(4 * (B + 1)) + 2 * a256(4 * B * L + 4) + a256(8 * B)
    lse = ws[off:off + 4 * B * T].view(torch.float32)
    dbg = lse[B * T - 64:B * T - 32].cpu().numpy().reshape(4, 2, 4)
    for b in range(4):
        for d in range(2):
            c1, cb, c2, n1 = dbg[b, d]
            n2 = int(act_lens[b].item()) - int(n1)
            print("b=%d %s: first half %d steps %.0f cyc/step | barrier %.0f cyc | second half %d steps %.0f cyc/step" %
                  (b, "alpha" if d == 0 else "beta ", n1, c1 / max(n1, 1), cb, n2, c2 / max(n2, 1)))


if __name__ == "__main__":
    main()
