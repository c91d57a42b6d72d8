"""Oracle (test infrastructure): greedy CTC decode, restating ArgmaxDecoder.decode
(reference src/decoder.py:116-185, identical to CnnOcrModel.decode_without_lm, src/models/cnnlstm.py:479-541).

Pinned against the reference's own decoder, imported from /root/reference, by tests/golden/make_golden.py
(fixtures tests/golden/decode_*.npz) and tests/test_oracle_vs_reference.py.
"""
import numpy as np


def uxxxx_to_utf8(in_str):
    # reference src/textutils.py:216-243
    if in_str.strip() == "":
        return ""
    out = ""
    for tok in in_str.split():
        out += tok if tok in ("<unk>", "<s>", "</s>") else chr(int(tok[1:], 16))
    return out


def decode_loop(logits, lens, idx_to_char, uxxxx=True):
    """Literal frame-by-frame restatement (decoder.py:130-185).  logits: float32 ndarray [T,B,A]."""
    logits = np.asarray(logits, dtype=np.float32)
    T, B, A = logits.shape
    thresh = 3 * 1 / len(idx_to_char)  # python float; NumPy 2 compares float32 < float in float32 (NEP 50)
    prev = [""] * B
    res = [""] * B
    for t in range(T):
        frame = logits[t]
        mx = frame.max(1).flatten()
        am = frame.argmax(1).flatten()
        for b in range(B):
            if t >= lens[b]:
                continue
            if am[b] == 0:
                prev[b] = ""
                continue
            if mx[b] < thresh:
                prev[b] = ""
                continue
            ch = idx_to_char[int(am[b])]
            if prev[b] == ch:
                continue
            res[b] += ch
            prev[b] = ch
            if t != T - 1:
                res[b] += " "
    for b in range(B):
        if len(res[b]) > 0 and res[b][-1] == " ":
            res[b] = res[b][:-1]
    if not uxxxx:
        res = [uxxxx_to_utf8(r) for r in res]
    return res


def frame_path(logits, lens, n_symbols):
    """Per-frame label path [B,T] int32: argmax, 0 for blank / low confidence, -1 beyond lens[b]
    (the integer form of src/utils/visualization.py:111-157)."""
    logits = np.asarray(logits, dtype=np.float32)
    T, B, A = logits.shape
    thresh = np.float32(3 * 1 / n_symbols)
    if T == 0:
        return np.zeros((B, 0), np.int32)
    am = logits.argmax(2).astype(np.int32)  # [T,B], first max; NaN counts as max (numpy)
    mx = logits.max(2)
    lab = np.where((am == 0) | (mx < thresh), 0, am).astype(np.int32)
    tt = np.arange(T)[:, None]
    lab = np.where(tt < np.asarray(lens)[None, :], lab, -1)
    return np.ascontiguousarray(lab.T)


def collapse(path_row, canon=None):
    """Collapse repeats / drop blanks on one line's frame path -> list of label indices."""
    out = []
    prev = 0
    for v in path_row:
        if v < 0:
            continue
        c = 0 if v == 0 else (int(canon[v]) if canon is not None else int(v))
        if v > 0 and c != prev:
            out.append(int(v))
        prev = c
    return out


def decode_labels(logits, lens, n_symbols, canon=None):
    p = frame_path(logits, lens, n_symbols)
    return [collapse(row, canon) for row in p], p
