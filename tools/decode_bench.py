"""cfg5: batch decode throughput - synthetic mixed-width lines through CNN + BiLSTM + greedy decode, sharded by width
bucket over the ranks of one box (no data-path collective: lines are independent).  Widths follow the reference's
bucket mix at line height 30 (src/data/madcat.py:58-66 as quoted in SURVEY.md section 8d).  Each batch is uploaded from
pinned host memory, run through model.eval() forward and decoded to strings (decode_without_lm) inside the timed region.

    python tools/decode_bench.py [--lines 8192] [--batch 256]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/decode_bench.py
Prints one JSON line (rank 0): whole-job lines/s = lines decoded by all ranks / max-over-ranks time."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vistaocr_b200 import Alphabet, CnnOcrModel  # noqa: E402
from vistaocr_b200.sharding import shard_batches  # noqa: E402

MIX = [(0.10, 60, 150), (0.10, 150, 200), (0.25, 200, 300), (0.25, 300, 350), (0.20, 350, 450), (0.09, 450, 600),
       (0.01, 600, 1200)]


def synth_widths(rng, n):
    u = rng.random(n)
    edges = np.cumsum([m[0] for m in MIX])
    which = np.minimum(np.searchsorted(edges, u), len(MIX) - 1)
    lo = np.array([MIX[k][1] for k in which])
    hi = np.array([MIX[k][2] for k in which])
    return (lo + (rng.random(n) * (hi - lo))).astype(np.int32)


def run(lines, batch, dev, rank=0, world=1, height=30, n_symbols=120):
    torch.manual_seed(7)
    alphabet = Alphabet(["<ctc-blank>"] + ["u%04x" % (0x21 + i) for i in range(n_symbols - 1)])
    model = CnnOcrModel(alphabet=alphabet, verbose=False, input_line_height=height, rds_line_height=30, lstm_input_dim=128,
                        num_lstm_layers=3, num_lstm_hidden_units=512, p_lstm_dropout=0.5)
    model.eval()
    rng = np.random.default_rng(7)
    widths = synth_widths(rng, lines)
    plan = shard_batches(widths, height, batch, world, rank, drop_last=False)
    host = []
    for idx in plan:  # pre-built pinned host batches (image decoding / resizing is outside this benchmark)
        w = widths[idx]
        x = torch.zeros((len(idx), 1, height, int(w[0])), dtype=torch.float32)
        for b, wb in enumerate(w):
            x[b, :, :, :wb] = torch.from_numpy(rng.random((1, height, int(wb)), dtype=np.float32))
        host.append((x.pin_memory(), torch.from_numpy(w.astype(np.int32))))

    def one(xb, wb):
        with torch.no_grad():
            logits, lens = model(xb.to(dev, non_blocking=True), wb)
            return model.decode_without_lm(logits, lens, uxxxx=True)

    for xb, wb in host[:2]:
        one(xb, wb)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 0
    for xb, wb in host:
        n += len(one(xb, wb))
    torch.cuda.synchronize()
    return n, time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lines", type=int, default=8192)
    ap.add_argument("--batch", type=int, default=256)
    args = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    n, dt = run(args.lines * world, args.batch, dev, rank, world)
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([dt, float(n)], dtype=torch.float64, device=dev)
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dt, n = float(tmax[0]), int(t[1])
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps({"metric": "greedy-decode lines/sec", "value": n / dt, "unit": "lines/s", "n_gpus": world,
                          "lines": n, "batch": args.batch, "seconds": dt, "scaling": "weak",
                          "config": "cfg5: mixed-width lines (reference bucket mix, height 30), alphabet 120, H2D + eval "
                                    "forward + greedy decode to strings per batch, width-bucket sharding, no collective"}))


if __name__ == "__main__":
    main()
