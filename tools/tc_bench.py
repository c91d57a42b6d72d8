"""Tensor-core GEMM / convolution kernels alone at the shapes of a cfg2 training step (T*B = 18560 frames, H = 512;
conv stack at 30 x 594 / 15 x 415 / 7 x 290, batch 64): time per launch and useful TFLOP/s (one product's worth; the
fp32-contract mode issues three).  VOCR_TC_CLUSTER=0 runs the same kernels without thread-block clusters.

    python tools/tc_bench.py [gemm|conv|all]
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vistaocr_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def t(fn, iters=7):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()  # > L2
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e-3)
    return float(np.median(ts))


def gemms():
    TB = 18560
    shapes = {"xproj fwd l1/l2 (K,K)": (0, 1, TB, 4096, 1024), "xproj fwd l0 (K,K)": (0, 1, TB, 4096, 128),
              "dx l1/l2 (K,MN)": (0, 0, TB, 1024, 4096), "dW_ih (MN,MN)": (1, 0, 4096, 1024, TB),
              "dW_hh (MN,MN)": (1, 0, 2048, 512, TB), "bridge fwd": (0, 1, TB, 128, 1792),
              "prob fwd": (0, 1, TB, 96, 1024)}
    for name, (ta, tb, M, N, K) in shapes.items():
        A = torch.randn((K, M) if ta else (M, K), device=dev)
        Bm = torch.randn((N, K) if tb else (K, N), device=dev)
        C = torch.empty((M, N), device=dev)
        Ao, Bo = ops.Operand(A), ops.Operand(Bm)
        Ao.split16()
        Bo.split16()
        flop = 2.0 * M * N * K
        sec = t(lambda: ops.mm(ta, tb, M, N, K, Ao, A.shape[1], Bo, Bm.shape[1], C, N))
        ref = (A.double().t() if ta else A.double()) @ (Bm.double().t() if tb else Bm.double())
        err = ((C.double() - ref).abs().max() / ref.abs().max()).item()
        print(json.dumps({"gemm": name, "M": M, "N": N, "K": K, "us": round(sec * 1e6, 1),
                          "useful_TFLOPs": round(flop / sec / 1e12, 1), "rel_err": err}), flush=True)


def convs():
    for name, (B, H, W, Cin, Cout) in {"conv 64->64 @30x594": (64, 30, 594, 64, 64), "conv 64->128 @15x415": (64, 15, 415, 64, 128),
                                       "conv 128->128 @15x415": (64, 15, 415, 128, 128),
                                       "conv 128->256 @7x290": (64, 7, 290, 128, 256),
                                       "conv 256->256 @7x290": (64, 7, 290, 256, 256)}.items():
        x = torch.randn(B, H, W, Cin, device=dev)
        w = torch.randn(Cout, Cin, 3, 3, device=dev) * 0.05
        b = torch.zeros(Cout, device=dev)
        flop = 2.0 * B * H * W * 9 * Cin * Cout
        xo = ops.Operand(x)
        xo.split16()
        sec = t(lambda: ops.conv3x3(x, w, b, x_op=xo))
        dz = torch.randn(B, H, W, Cout, device=dev)
        dzo = ops.Operand(dz)
        dzo.split16()
        rec = {"conv": name, "fwd_us": round(sec * 1e6, 1), "fwd_useful_TFLOPs": round(flop / sec / 1e12, 1)}
        sec = t(lambda: ops.conv3x3_dgrad(dz, w, dz_op=dzo))
        rec.update(dgrad_us=round(sec * 1e6, 1), dgrad_useful_TFLOPs=round(flop / sec / 1e12, 1))
        sec = t(lambda: ops.conv3x3_wgrad(x, dz, x_op=xo, dz_op=dzo))
        rec.update(wgrad_us=round(sec * 1e6, 1), wgrad_useful_TFLOPs=round(flop / sec / 1e12, 1))
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    print(json.dumps({"cluster": os.environ.get("VOCR_TC_CLUSTER", "1"), "precision": ops.get_precision()}))
    if what in ("gemm", "all"):
        gemms()
    if what in ("conv", "all"):
        convs()
