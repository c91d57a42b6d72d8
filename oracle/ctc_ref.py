"""Oracle (test infrastructure): python face of oracle/ctc_ref.c plus a brute-force path enumerator.
See ctc_ref.c for what is restated and how parity is pinned."""
import ctypes
import itertools

import numpy as np

from .build_oracle import build

_libs = {}


def _lib(real):
    if real not in _libs:
        l = ctypes.CDLL(build()[real])
        l.ctc_ref.restype = ctypes.c_int
        l.ctc_ref.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int] * 3 + [ctypes.c_void_p]
        _libs[real] = l
    return _libs[real]


def ctc_ref(acts, labels, act_lens, label_lens, want_grads=True, real="double"):
    """acts float32 [T,B,A]; returns (costs float64[B], grads float32[T,B,A] or None)."""
    acts = np.ascontiguousarray(acts, dtype=np.float32)
    T, B, A = acts.shape
    labels = np.ascontiguousarray(labels, dtype=np.int32)
    act_lens = np.ascontiguousarray(act_lens, dtype=np.int32)
    label_lens = np.ascontiguousarray(label_lens, dtype=np.int32)
    costs = np.zeros(B, np.float64)
    grads = np.zeros_like(acts) if want_grads else None
    dummy = np.zeros(1, np.int32)
    lab = labels if labels.size else dummy
    _lib(real).ctc_ref(acts.ctypes.data, grads.ctypes.data if want_grads else None, lab.ctypes.data,
                       label_lens.ctypes.data, act_lens.ctypes.data, T, B, A, costs.ctypes.data)
    return costs, grads


def ctc_brute_force(acts_tA, label):
    """-ln sum over ALL A^T frame paths that collapse to `label` (tiny T only).  acts_tA: [T,A] float."""
    x = np.asarray(acts_tA, dtype=np.float64)
    T, A = x.shape
    lp = x - np.log(np.exp(x - x.max(1, keepdims=True)).sum(1, keepdims=True)) - x.max(1, keepdims=True)
    total = 0.0
    label = list(label)
    for path in itertools.product(range(A), repeat=T):
        out = []
        prev = -1
        for k in path:
            if k != prev and k != 0:
                out.append(k)
            prev = k
        if out == label:
            total += np.exp(sum(lp[t, k] for t, k in enumerate(path)))
    return -np.log(total) if total > 0 else np.inf
