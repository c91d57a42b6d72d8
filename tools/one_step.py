"""Minimal cfg2 training loop for profilers: WARM untimed steps, then STEPS steps on one resident batch (GPU box only).
ncu: every step ends with clamp_adam_kernel, which marks the step boundaries in a launch list."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from vistaocr_b200 import Alphabet, ClampAdam, CnnOcrModel, CTCLoss, train_step

dev = torch.device("cuda:0")
torch.manual_seed(7)
alphabet = Alphabet(["<ctc-blank>"] + ["u%04x" % (0x21 + i) for i in range(bench.TRAIN["train_cfg2"]["n_symbols"] - 1)])
model = CnnOcrModel(alphabet=alphabet, verbose=False, **bench.TRAIN["train_cfg2"]["hp"])
model.train()
crit, opt = CTCLoss(host_cost=False), ClampAdam(model.parameters(), lr=1e-3)
host = bench.synth_train_batches(bench.TRAIN["train_cfg2"], 1000, 1)
res = [(b[0].to(dev), b[1].to(dev), b[2], b[3], b[4]) for b in host]
for i in range(int(os.environ.get("WARM", 3)) + int(os.environ.get("STEPS", 1))):
    train_step(res[0], model, crit, opt)
torch.cuda.synchronize()
print("done")
