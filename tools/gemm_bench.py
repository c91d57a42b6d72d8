"""GEMM engines side by side at the LSTM-projection shapes of cfg2 (T*B = 18816, H = 512): tcgen05 3xTF32 vs FFMA."""
import json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vistaocr_b200 import ops

dev = torch.device("cuda:0")


def t(fn, iters=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e-3)
    return float(np.median(ts))


for name, (ta, tb, M, N, K) in {"xproj fwd (KK)": (0, 1, 18816, 4096, 1024), "dx (K,MN)": (0, 0, 18816, 1024, 4096),
                                "dW_ih (MN,MN)": (1, 0, 4096, 1024, 18816), "dW_hh (MN,MN)": (1, 0, 2048, 512, 18752),
                                "bridge fwd": (0, 1, 18816, 128, 1792), "prob fwd": (0, 1, 18816, 96, 1024)}.items():
    A = torch.randn((K, M) if ta else (M, K), device=dev)
    Bm = torch.randn((N, K) if tb else (K, N), device=dev)
    C = torch.empty((M, N), device=dev)
    Ao, Bo = ops.Operand(A), ops.Operand(Bm)
    Ao.split(); Bo.split()
    flop = 2.0 * M * N * K
    t_tc = t(lambda: ops.tc_gemm(ta, 0 if tb else 1, M, N, K, Ao.split(), A.shape[1], Bo.split(), Bm.shape[1], C, N))
    t_ff = t(lambda: ops.gemm(ta, tb, M, N, K, A, A.shape[1], Bm, Bm.shape[1], C, N))
    t_sp = t(lambda: ops.split_tf32(A))
    print(json.dumps({"gemm": name, "M": M, "N": N, "K": K, "tc_ms": t_tc * 1e3, "tc_TFLOPs": flop / t_tc / 1e12,
                      "ffma_ms": t_ff * 1e3, "ffma_TFLOPs": flop / t_ff / 1e12, "split_A_ms": t_sp * 1e3}), flush=True)
