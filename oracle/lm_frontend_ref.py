"""Oracle (test infrastructure): the host half of LmDecoder.decode restated (reference src/decoder.py:36-59,61-101).
Pinned against the reference method itself (object built without __init__, capturing executor) in
tests/test_oracle_vs_reference.py."""
import numpy as np
import torch


def lm_remap_ref(model_output, lens, idx_to_char, lm_units):
    """model_output: float32 tensor [T,B,A] -> list of float64 arrays [len_b, 1+len(lm_units)]."""
    units = ["<ctc-blank>"] + list(lm_units)
    lmchar_to_idx = dict(zip(units, range(len(units))))
    model_idx, lm_idx = [], []
    for m in range(len(idx_to_char)):
        ch = idx_to_char[m]
        if ch in lmchar_to_idx:
            model_idx.append(m)
            lm_idx.append(lmchar_to_idx[ch])
    T, B, A = model_output.shape
    probs = torch.nn.functional.log_softmax(model_output.reshape(-1, A), dim=1).view(T, B, -1).cpu()
    out = []
    for b in range(B):
        r = np.full((int(lens[b]), len(units)), np.log(1e-10))
        r[:, lm_idx] = probs[:int(lens[b]), b, model_idx]
        out.append(r)
    return out
