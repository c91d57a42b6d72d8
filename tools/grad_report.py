"""Diagnostic: per-parameter gradient error of the CUDA path vs the float64 oracle, next to the error of the
reference's own fp32 run (golden fixture) - shows which tensors are intrinsically ill-conditioned in fp32."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import model_ref as M  # noqa: E402
from tests.test_gpu_model import CONFIGS, _model  # noqa: E402
from vistaocr_b200 import CTCLoss  # noqa: E402

dev = torch.device("cuda:0")
for name in ("h30", "h60", "h120"):
    hp = CONFIGS[name]
    z = np.load(os.path.join("tests", "golden", "model_%s.npz" % name))
    A = int(z["n_symbols"])
    sd = M.make_state_dict(hp, A, seed=int(z["seed"]))
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    for k, v in sd64.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    u1, u2 = torch.from_numpy(z["u1"]), torch.from_numpy(z["u2"])
    want, wl = M.forward_ref(sd64, torch.from_numpy(z["x"]).double(), z["widths"], hp, (u1, u2), training=True,
                             bn_updates={}, use_nn_lstm=False)
    wloss = M.ctc_sum_ref(want, z["labels"], wl, z["label_lens"])
    wloss.backward()
    model = _model(hp, A, sd, dev)
    model.cnn[6]._random_samples, model.cnn[13]._random_samples = u1, u2
    model.train()
    logits, lens = model(torch.from_numpy(z["x"]).to(dev), torch.from_numpy(z["widths"]))
    loss = CTCLoss()(logits, torch.from_numpy(z["labels"]), lens, torch.from_numpy(z["label_lens"]))
    loss.backward()
    print(name, "loss ours %.6f  f64 %.6f | logits err %.2e" % (loss.item(), wloss.item(),
          (logits.detach().double().cpu() - want.detach()).abs().max().item()))
    for k, p in model.named_parameters():
        g64 = sd64[k].grad
        e = (p.grad.double().cpu() - g64).abs().max().item() / max(g64.abs().max().item(), 1e-30)
        ref = ""
        if "grad." + k in z.files:
            ref = "  reference-fp32 rel err %.2e" % (np.abs(z["grad." + k] - g64.numpy()).max() / np.abs(g64.numpy()).max())
        print("   %-28s max|g| %9.3g  ours rel err %.2e%s" % (k, g64.abs().max().item(), e, ref))
