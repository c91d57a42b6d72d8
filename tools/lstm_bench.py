"""BiLSTM layer recurrence at cfg2 size (T=294, B=64, H=512): device time of the persistent fwd / bwd kernels."""
import json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vistaocr_b200 import _lib, ops

dev = torch.device("cuda:0")
T, B, D, H = int(os.environ.get("T", 294)), int(os.environ.get("B", 64)), 1024, 512
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn((T, B, D), device=dev, generator=g).requires_grad_(True)
k = 1.0 / H ** 0.5
w_ih = ((torch.rand((8 * H, D), device=dev, generator=g) * 2 - 1) * k).requires_grad_(True)
w_hh = ((torch.rand((2, 4 * H, H), device=dev, generator=g) * 2 - 1) * k).requires_grad_(True)
bias = ((torch.rand((8 * H,), device=dev, generator=g) * 2 - 1) * k).requires_grad_(True)
lens = torch.from_numpy(np.sort(np.random.default_rng(0).integers(T // 2, T + 1, size=B))[::-1].astype(np.int32).copy())
lens[0] = T
lens_dev = lens.to(dev)
dy = torch.randn((T, B, 2 * H), device=dev, generator=g)
iters = int(os.environ.get("ITERS", 3))
for it in range(iters):
    _lib.PROFILER.reset()
    _lib.PROFILER.timing = True
    y = ops.bilstm_layer(x, w_ih, w_hh, bias, lens_dev, T)
    y.backward(dy)
    torch.cuda.synchronize()
    s = _lib.PROFILER.summary()
    _lib.PROFILER.timing = False
print(json.dumps({k_: round(v["ms"], 3) for k_, v in s.items()}))
print(json.dumps({"T": T, "B": B, "fwd_us_per_step": s["vocr_bilstm_fwd_f32"]["ms"] * 1e3 / T,
                  "bwd_us_per_step": s["vocr_bilstm_bwd_f32"]["ms"] * 1e3 / T}))
