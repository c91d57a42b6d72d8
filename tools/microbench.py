"""Kernel-level timings (CUDA events, L2 flushed between iterations) for the HBM-bound kernels:
greedy decode and CTC fwd+bwd.  Prints one JSON object per case.  Not the contract bench (see bench.py)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vistaocr_b200.decoder import greedy_decode_labels  # noqa: E402
from vistaocr_b200.warpctc import ctc_costs_and_grads  # noqa: E402


def peak_hbm():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured"
    return 6650.0, "fallback"


def time_it(fn, iters=10, warmup=3, flush=None):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.add_(1.0)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e-3)
    return float(np.median(ts)), float(np.min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="decode,ctc")
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    flush = torch.zeros(256 << 20, dtype=torch.float32, device=dev)  # 1 GiB > 126 MB L2
    peak, how = peak_hbm()
    g = torch.Generator(device="cuda").manual_seed(7)
    if "decode" in args.what:
        for (T, B, A) in [(392, 64, 120), (245, 512, 120), (392, 2048, 120), (244, 2048, 166)]:
            x = torch.randn((T, B, A), device=dev, generator=g)
            lens = torch.randint(T // 2, T + 1, (B,), device=dev, generator=g, dtype=torch.int32)
            med, best = time_it(lambda: greedy_decode_labels(x, lens, 3 / A), args.iters, flush=flush)
            byt = T * B * A * 4
            print(json.dumps({"kernel": "greedy_decode", "T": T, "B": B, "A": A, "ms": med * 1e3, "best_ms": best * 1e3,
                              "GBs": byt / med / 1e9, "frac_hbm": byt / med / 1e9 / peak, "peak": how,
                              "lines_per_s": B / med}), flush=True)
    if "ctc" in args.what:
        rng = np.random.default_rng(4)
        for (T, B, A, L) in [(100, 256, 200, 20), (250, 256, 120, 50), (500, 256, 120, 50), (1000, 256, 200, 150),
                             (1000, 256, 80, 20), (290, 64, 96, 40), (100, 2048, 200, 20)]:
            x = torch.randn((T, B, A), device=dev, generator=g)
            act_lens = torch.from_numpy(np.sort(rng.integers(T // 2, T + 1, size=B))[::-1].astype(np.int32).copy())
            label_lens = torch.full((B,), L, dtype=torch.int32)
            labels = torch.from_numpy(rng.integers(1, A, size=B * L).astype(np.int32))
            ld, al, ll = labels.to(dev), act_lens.to(dev), label_lens  # label_lens stays on host (sizes the workspace)
            med, best = time_it(lambda: ctc_costs_and_grads(x, ld, al, ll), args.iters, flush=flush)
            byt = 2 * T * B * A * 4
            print(json.dumps({"kernel": "ctc_fwd_bwd", "T": T, "B": B, "A": A, "L": L, "ms": med * 1e3,
                              "best_ms": best * 1e3, "GBs": byt / med / 1e9, "frac_hbm": byt / med / 1e9 / peak,
                              "peak": how}), flush=True)


if __name__ == "__main__":
    main()
