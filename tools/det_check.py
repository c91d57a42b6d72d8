import sys, os
sys.path.insert(0, os.getcwd())
import torch, numpy as np
from tests.test_gpu_graphs import _model, _batches
from vistaocr_b200 import ClampAdam, CTCLoss, GraphedTrainStep, train_step
seq = _batches()
def run(mode):
    m = _model(0.0); m.train()
    o = ClampAdam(m.parameters(), lr=1e-2); c = CTCLoss(host_cost=False)
    st = GraphedTrainStep(m, c, o, capture_after=1)
    out = []
    for b in seq:
        out.append((train_step(b, m, c, o) if mode == "eager" else st(b))[0].item())
    return out
a = run("eager"); b = run("eager"); g = run("graph")
print("eager1", a); print("eager2", b); print("graph ", g)
print("eager deterministic:", a == b, " graph == eager:", a == g)
