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
    ap.add_argument("--what", default="decode,ctc,preproc")
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
    if "preproc" in args.what:
        from vistaocr_b200.imagetransforms import scaled_width  # noqa: E402
        from vistaocr_b200 import _lib  # noqa: E402
        rng = np.random.default_rng(9)
        for (h, H, B, wlo, whi) in [(90, 30, 512, 600, 2400), (120, 60, 256, 800, 2400), (180, 120, 128, 600, 3000)]:
            ws_ = rng.integers(wlo, whi + 1, size=B).astype(np.int32)
            hs_ = np.full(B, h, np.int32)
            dws = np.array([scaled_width(h, int(w), H) for w in ws_], np.int32)
            offs = np.zeros(B, np.int64)
            offs[1:] = np.cumsum(hs_[:-1].astype(np.int64) * ws_[:-1])
            n_in = int((hs_.astype(np.int64) * ws_).sum())
            d_pix = torch.randint(0, 256, (n_in,), dtype=torch.uint8, device=dev)
            d_meta = torch.from_numpy(np.concatenate([hs_, ws_, dws])).to(dev)
            d_offs = torch.from_numpy(offs).to(dev)
            w_out = int(dws.max())
            out = torch.empty((B, 1, H, w_out), dtype=torch.float32, device=dev)
            l = _lib.lib()

            def run():
                _lib.check(l.vocr_scale_lines_u8(_lib.ptr(d_pix), _lib.ptr(d_offs), _lib.ptr(d_meta[0:B]),
                                                 _lib.ptr(d_meta[B:2 * B]), _lib.ptr(d_meta[2 * B:3 * B]), None, B, 1, H,
                                                 w_out, 1, 15, _lib.ptr(out), _lib.stream()), "scale")
            med, best = time_it(run, args.iters, flush=flush)
            byt = n_in + out.numel() * 4
            print(json.dumps({"kernel": "scale_lines_u8", "h": h, "H": H, "B": B, "Wout": w_out, "ms": med * 1e3,
                              "best_ms": best * 1e3, "GBs": byt / med / 1e9, "frac_hbm": byt / med / 1e9 / peak,
                              "peak": how, "lines_per_s": B / med}), flush=True)


if __name__ == "__main__":
    main()
