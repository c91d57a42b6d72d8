"""Measured parity errors of the CUDA path (GPU box only) -> gpurun_out/parity_report.json, the source of
profiles/r02_parity_errors.md.

Part 1: N seeds of a small lively model (dropout 0.5 with injected masks): per parameter tensor the max-abs error of the
gradient relative to the tensor's max, against the float64 oracle - for the CUDA path ("ours") and for the oracle run
in float32 on the CPU ("ref32", the reference's own arithmetic).  For conv weights the share of the squared error that
sits in ONE output channel is reported too: ~1.0 is the signature of a single ReLU / max-pool decision that flipped
between two fp32 evaluations (all of that channel's weights see one term more or less), not of an arithmetic error.
Part 2: the full-size records written by tests/test_gpu_model.py (gpurun_out/parity_errors.jsonl) are summarised.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import model_ref as M  # noqa: E402
from vistaocr_b200 import Alphabet, CnnOcrModel, CTCLoss  # noqa: E402

dev = torch.device("cuda:0")
hp = dict(input_line_height=30, rds_line_height=30, lstm_input_dim=32, num_lstm_layers=3, num_lstm_hidden_units=40,
          p_lstm_dropout=0.5)
A, B = 23, 5
N = int(os.environ.get("SEEDS", 12))
worst = {}
for seed in range(N):
    sd = M.make_state_dict(hp, A, seed=100 + seed)
    model = CnnOcrModel(alphabet=Alphabet(["<ctc-blank>"] + ["u%04x" % (0x61 + i) for i in range(A - 1)]),
                        verbose=False, **hp)
    model.load_state_dict(sd, strict=True)
    rng = np.random.default_rng(seed)
    x, widths, labels, label_lens = M.synth_batch(rng, B, 30, 40, 170, A, 2, 10)
    u1 = torch.from_numpy(rng.random((B, 64, 2)).astype(np.float32))
    u2 = torch.from_numpy(rng.random((B, 128, 2)).astype(np.float32))
    model.cnn[6]._random_samples, model.cnn[13]._random_samples = u1, u2
    lens = [M.out_hw(30, int(w), 0)[1] for w in widths]
    tmax, wf, H2 = max(lens), M.out_hw(30, int(widths[0]), 0)[1], 2 * hp["num_lstm_hidden_units"]
    keep = [(rng.random((tmax, B, H2)) < 0.5).astype(np.uint8) for _ in range(2)]
    model._dropout_masks = [torch.from_numpy(k) for k in keep]
    model.train()
    logits, olens = model(torch.from_numpy(x).to(dev), torch.from_numpy(widths))
    CTCLoss()(logits, torch.from_numpy(labels), olens, torch.from_numpy(label_lens)).backward()

    def oracle(dtype):
        s = {k: (v.to(dtype).clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else
                 (v.to(dtype) if v.is_floating_point() else v)) for k, v in sd.items()}
        masks = []
        for k in keep:
            m = torch.ones((wf, B, H2), dtype=dtype)
            m[:tmax] = torch.from_numpy(k).to(dtype) * 2.0
            masks.append(m)
        out, ol = M.forward_ref(s, torch.from_numpy(x).to(dtype), widths, hp, (u1, u2), training=True, bn_updates={},
                                dropout_masks=masks, use_nn_lstm=False)
        M.ctc_sum_ref(out, labels, ol, label_lens).backward()
        return out.detach(), s

    w64, s64 = oracle(torch.float64)
    w32, s32 = oracle(torch.float32)
    scale = w64.abs().max().item()
    rec = worst.setdefault("logits", {"ours": 0.0, "ref32": 0.0})
    rec["ours"] = max(rec["ours"], (logits.detach().double().cpu() - w64).abs().max().item() / scale)
    rec["ref32"] = max(rec["ref32"], (w32.double() - w64).abs().max().item() / scale)
    for k, p in model.named_parameters():
        if k.startswith("cnn.") and k.endswith(".bias") and int(k.split(".")[1]) in M.CONV_IDX:
            continue
        g64 = s64[k].grad
        gs = g64.abs().max().item()
        d = p.grad.double().cpu() - g64
        d32 = s32[k].grad.double() - g64
        rec = worst.setdefault(k, {"ours": 0.0, "ref32": 0.0, "ours_l2": 0.0, "ref32_l2": 0.0, "ours_top_channel_share": None,
                                   "ref32_top_channel_share": None})
        e, e32 = d.abs().max().item() / gs, d32.abs().max().item() / gs
        if e >= rec["ours"]:
            rec["ours"] = e
            if d.dim() == 4:
                per = (d ** 2).sum(dim=(1, 2, 3))
                rec["ours_top_channel_share"] = (per.max() / per.sum().clamp_min(1e-300)).item()
        if e32 >= rec["ref32"]:
            rec["ref32"] = e32
            if d32.dim() == 4:
                per = (d32 ** 2).sum(dim=(1, 2, 3))
                rec["ref32_top_channel_share"] = (per.max() / per.sum().clamp_min(1e-300)).item()
        rec["ours_l2"] = max(rec["ours_l2"], (d.norm() / g64.norm()).item())
        rec["ref32_l2"] = max(rec["ref32_l2"], (d32.norm() / g64.norm()).item())
    del model

out = {"small_model": {"seeds": N, "config": hp, "worst_over_seeds": worst}, "full_size": []}
p = os.path.join(ROOT, "gpurun_out", "parity_errors.jsonl")
if os.path.exists(p):
    for line in open(p):
        r = json.loads(line)
        g = r.pop("grads")
        up = [v for k, v in g.items() if k.startswith("rapid_ds.") or (k.startswith("cnn.") and int(k.split(".")[1]) <= 11)]
        dn = [v for k, v in g.items() if not (k.startswith("rapid_ds.") or (k.startswith("cnn.") and int(k.split(".")[1]) <= 11))]
        r["grad_upstream_of_pool"] = {"ours_max": max(v["ours"] for v in up), "ref32_max": max(v["ref32"] for v in up)}
        r["grad_downstream"] = {"ours_max": max(v["ours"] for v in dn), "ref32_max": max(v["ref32"] for v in dn)}
        r["grad_worst_tensors"] = sorted(((k, v["ours"], v["ref32"]) for k, v in g.items()), key=lambda t: -t[1])[:6]
        out["full_size"].append(r)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "parity_report.json"), "w"), indent=1)
print(json.dumps(out["small_model"]["worst_over_seeds"], indent=1)[:6000])
for r in out["full_size"]:
    print(json.dumps(r))
