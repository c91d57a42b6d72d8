"""Validation scoring on the device: CER / WER of a batch of hypotheses, with the O(n*m) edit
distance of every line computed by one CUDA kernel launch instead of a NumPy double loop per line
(reference src/textutils.py:264-351 `edit_distance`, `form_tokenized_words`, `compute_cer_wer`, used by
`test_on_val`, src/train_cnn_lstm.py:61-79).  Tokenisation (splitting uxxxx strings, grouping words, punctuation
and digits) stays on the host exactly as the reference does it; only integer ids go to the GPU.
"""
import numpy as np
import torch

from . import _lib
from ._lib import check, lib, ptr, stream

_PUNCT = {"u002e", "u002c", "u003b", "u0027", "u0022", "u002f", "u0021", "u0028", "u0029", "u005b", "u005d", "u003c",
          "u003e", "u002d", "u005f", "u007b", "u007d", "u0024", "u0025", "u0023", "u0026", "u060c", "u201d", "u060d",
          "u060f", "u061f", "u066d", "ufd3e", "ufd3f", "u061e", "u066a", "u066b", "u066c", "u002a", "u002b", "u003a",
          "u003d", "u005e", "u0060", "u007c", "u007e"}
_DIGITS = {"u0660", "u0661", "u0662", "u0663", "u0664", "u0665", "u0666", "u0667", "u0668", "u0669", "u0030", "u0031",
           "u0032", "u0033", "u0034", "u0035", "u0036", "u0037", "u0038", "u0039"}


def form_tokenized_words(chars, with_spaces=False):
    """reference textutils.py:290-323: words are '_'-joined runs of characters; u0020 separates words; punctuation and
    digits are words of their own."""
    words = []
    start = 0
    for i, ch in enumerate(chars):
        if ch == "u0020":
            if start != i:
                words.append("_".join(chars[start:i]))
                if with_spaces:
                    words.append("u0020")
            start = i + 1
            continue
        if ch in _PUNCT or ch in _DIGITS:
            if start != i:
                words.append("_".join(chars[start:i]))
            words.append(ch)
            start = i + 1
            continue
        if i == len(chars) - 1:
            words.append(chars[start] if start == i else "_".join(chars[start:]))
    return words


def _strip_spaces(words):
    while len(words) > 0 and words[0] == "u0020":
        words = words[1:]
    while len(words) > 0 and words[-1] == "u0020":
        words = words[:-1]
    return words


def edit_distances(a_seqs, b_seqs, device="cuda"):
    """Edit distance of P pairs of int sequences (lists / arrays) -> int32 CPU tensor [P]."""
    P = len(a_seqs)
    assert P == len(b_seqs)
    if P == 0:
        return torch.zeros(0, dtype=torch.int32)
    a_off = np.zeros(P + 1, np.int32)
    b_off = np.zeros(P + 1, np.int32)
    a_off[1:] = np.cumsum([len(s) for s in a_seqs])
    b_off[1:] = np.cumsum([len(s) for s in b_seqs])
    cat = lambda seqs, n: (np.concatenate([np.asarray(s, np.int32).reshape(-1) for s in seqs]) if n else
                           np.zeros(1, np.int32))
    dev = torch.device(device)
    d = [torch.from_numpy(x).to(dev, non_blocking=True) for x in (cat(a_seqs, a_off[-1]), a_off, cat(b_seqs, b_off[-1]),
                                                                   b_off)]
    dist = torch.empty((P,), dtype=torch.int32, device=dev)
    max_n = int(np.diff(a_off).max())
    max_m = int(np.diff(b_off).max())
    st = lib().vocr_edit_distance_i32(ptr(d[0]), ptr(d[1]), ptr(d[2]), ptr(d[3]), P, max_n, max_m, ptr(dist), stream())
    check(st, "vocr_edit_distance_i32")
    return dist.cpu()


def compute_cer_wer_batch(hyp_transcriptions, ref_transcriptions, device="cuda"):
    """Per-line (cer, wer) exactly as the reference's compute_cer_wer(hyp, ref) returns them, for a whole batch."""
    ids = {}
    tok = lambda t: ids.setdefault(t, len(ids))
    hc, rc, hw, rw = [], [], [], []
    for hyp, ref in zip(hyp_transcriptions, ref_transcriptions):
        hyp_chars, ref_chars = hyp.split(" "), ref.split(" ")
        hc.append([tok(c) for c in hyp_chars])
        rc.append([tok(c) for c in ref_chars])
        hw.append([tok("w:" + w) for w in _strip_spaces(form_tokenized_words(hyp_chars))])
        rw.append([tok("w:" + w) for w in _strip_spaces(form_tokenized_words(ref_chars))])
    d = edit_distances(hc + hw, rc + rw, device).tolist()
    n = len(hc)
    out = []
    for i in range(n):
        out.append((float(d[i]) / len(rc[i]), float(d[n + i]) / len(rw[i])))  # ZeroDivisionError like the reference
    return out
