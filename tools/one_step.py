"""Minimal training loop for profilers: WARM untimed steps, then STEPS steps on one resident batch (GPU box only).
WORKLOAD=train_cfg2 (default) | train_cfg3; DECODE=1 runs eval forward + greedy decode of one cfg5-style batch instead.
ncu: every training step ends with clamp_adam_kernel, which marks the step boundaries in a launch list."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from vistaocr_b200 import Alphabet, ClampAdam, CnnOcrModel, CTCLoss, train_step

dev = torch.device("cuda:0")
torch.manual_seed(7)
n_iter = int(os.environ.get("WARM", 3)) + int(os.environ.get("STEPS", 1))
if os.environ.get("DECODE"):
    from vistaocr_b200 import GraphedDecoder
    model = bench.decode_model()
    dec = GraphedDecoder(model, capture_after=1 << 30)  # never captures: the fused inference tail, kernel by kernel
    x, w = bench.synth_decode_batches("decode_cfg5", 0, 1, int(os.environ.get("BATCH", 448)), max_batches=3)[1]
    x = x.to(dev)
    for i in range(n_iter):
        dec.labels(x, w, allow_eager=True)
else:
    import vistaocr_b200
    cfg = bench.TRAIN[os.environ.get("WORKLOAD", "train_cfg2")]
    vistaocr_b200.set_precision(cfg["precision"])
    alphabet = Alphabet(["<ctc-blank>"] + ["u%04x" % (0x21 + i) for i in range(cfg["n_symbols"] - 1)])
    model = CnnOcrModel(alphabet=alphabet, verbose=False, **cfg["hp"])
    model.train()
    crit, opt = CTCLoss(host_cost=False), ClampAdam(model.parameters(), lr=1e-3)
    host = bench.synth_train_batches(cfg, 1000, 1)
    res = [(b[0].to(dev), b[1].to(dev), b[2], b[3], b[4]) for b in host]
    for i in range(n_iter):
        train_step(res[0], model, crit, opt)
torch.cuda.synchronize()
print("done")
