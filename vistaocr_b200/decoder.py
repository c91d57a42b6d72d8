"""Greedy CTC decoder with the reference's `ArgmaxDecoder` interface (reference src/decoder.py:112-185).

`decode(model_output[T,B,A], batch_actual_timesteps[B], uxxxx=False, lang=None) -> list[str]`.
The per-frame argmax, blank / low-confidence mapping, repeat collapse and compaction run in two CUDA kernels
(csrc/decode.cu) behind `vocr_greedy_decode_f32`; the host only maps the compacted int32 label sequences to
strings.  CPU tensors (the reference decodes in a child process on CPU logits, decode_testset.py:86-106,166) are
uploaded first - there is no CPU code path.
"""
import numpy as np
import torch

from . import _lib
from .textutils import uxxxx_to_utf8


def greedy_decode_labels(model_output, lens, thresh, canon=None):
    """Device part.  Returns (labels[B,T] int32, counts[B] int32, path[B,T] int32) as CUDA tensors (no sync).

    thresh is compared in float32, as NumPy 2 does for `np.float32 < python float`."""
    _lib.require_cuda(model_output, "model_output", torch.float32)
    T, B, A = model_output.shape
    dev = model_output.device
    lens = torch.as_tensor(lens).to(device=dev, dtype=torch.int32, non_blocking=True).contiguous()
    if lens.numel() != B:
        raise _lib.VocrError("batch_actual_timesteps must have %d entries, got %d" % (B, lens.numel()))
    ld = max(T, 1)
    path = torch.empty((B, ld), dtype=torch.int32, device=dev)
    labels = torch.empty((B, ld), dtype=torch.int32, device=dev)
    counts = torch.empty((B,), dtype=torch.int32, device=dev)
    if canon is not None:
        canon = torch.as_tensor(canon).to(device=dev, dtype=torch.int32).contiguous()
    st = _lib.lib().vocr_greedy_decode_f32(_lib.ptr(model_output), T, B, A, _lib.ptr(lens), float(np.float32(thresh)),
                                           _lib.ptr(canon), _lib.ptr(path), _lib.ptr(labels), _lib.ptr(counts), ld,
                                           _lib.stream())
    _lib.check(st, "vocr_greedy_decode_f32")
    return labels, counts, path[:, :T]


def collapse_labels(path, lens, canon=None):
    """Second half of the greedy decode on a frame path [B,T] (see ops.linear_argmax): (labels[B,T], counts[B])."""
    _lib.require_cuda(path, "path", torch.int32)
    B, T = path.shape
    dev = path.device
    lens = torch.as_tensor(lens).to(device=dev, dtype=torch.int32, non_blocking=True).contiguous()
    labels = torch.empty((B, T), dtype=torch.int32, device=dev)
    counts = torch.empty((B,), dtype=torch.int32, device=dev)
    if canon is not None:
        canon = torch.as_tensor(canon).to(device=dev, dtype=torch.int32).contiguous()
    st = _lib.lib().vocr_ctc_collapse_i32(_lib.ptr(path), T, B, _lib.ptr(lens), _lib.ptr(canon), _lib.ptr(labels),
                                          _lib.ptr(counts), T, _lib.stream())
    _lib.check(st, "vocr_ctc_collapse_i32")
    return labels, counts


def _canon_map(alphabet):
    """Two alphabet indices carrying the same string collapse in the reference (it compares strings,
    decoder.py:166); returns None when all strings are distinct."""
    n = len(alphabet)
    first = {}
    canon = np.arange(n, dtype=np.int32)
    dup = False
    for i in range(n):
        c = alphabet.idx_to_char[i]
        if c in first:
            canon[i] = first[c]
            dup = True
        else:
            first[c] = i
    return canon if dup else None


def labels_to_strings(labels, counts, alphabet, uxxxx):
    labels = labels.cpu().numpy()
    counts = counts.cpu().numpy()
    idx_to_char = alphabet.idx_to_char
    out = []
    for b in range(labels.shape[0]):
        s = " ".join(idx_to_char[int(k)] for k in labels[b, :counts[b]])
        out.append(s if uxxxx else uxxxx_to_utf8(s))
    return out


class ArgmaxDecoder:
    def __init__(self, alphabet):
        self.alphabet = alphabet

    def decode(self, model_output, batch_actual_timesteps, uxxxx=False, lang=None):
        alphabet = self.alphabet if lang is None else self.alphabet[lang]
        min_prob_thresh = 3 * 1 / len(alphabet)
        if not model_output.is_cuda:
            model_output = model_output.cuda(non_blocking=True)
        model_output = model_output.detach().float().contiguous()
        labels, counts, _ = greedy_decode_labels(model_output, batch_actual_timesteps, min_prob_thresh,
                                                 _canon_map(alphabet))
        return labels_to_strings(labels, counts, alphabet, uxxxx)

    def decode_alignment(self, model_output, batch_actual_timesteps, lang=None):
        """Per-frame label path [B,T] int32 (blank/low-confidence -> 0, t >= len -> -1): the integer form of
        the reference's alignment spans (src/utils/visualization.py:111-157)."""
        alphabet = self.alphabet if lang is None else self.alphabet[lang]
        if not model_output.is_cuda:
            model_output = model_output.cuda(non_blocking=True)
        model_output = model_output.detach().float().contiguous()
        _, _, path = greedy_decode_labels(model_output, batch_actual_timesteps, 3 * 1 / len(alphabet), None)
        return path
