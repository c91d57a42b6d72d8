"""Per-call device times of one cfg2 training step, in launch order, with the integer arguments (shapes) of each
C-ABI call and the achieved rate (GPU box only)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from vistaocr_b200 import Alphabet, ClampAdam, CnnOcrModel, CTCLoss, _lib, train_step

dev = torch.device("cuda:0")
torch.manual_seed(7)
alphabet = Alphabet(["<ctc-blank>"] + ["u%04x" % (0x21 + i) for i in range(bench.N_SYMBOLS - 1)])
model = CnnOcrModel(alphabet=alphabet, verbose=False, **bench.CFG)
model.train()
crit, opt = CTCLoss(host_cost=False), ClampAdam(model.parameters(), lr=1e-3)
host = bench.synth_batches(1000, 1)
res = [(b[0].to(dev), b[1].to(dev), b[2], b[3], b[4]) for b in host]
for i in range(3):
    train_step(res[0], model, crit, opt)
torch.cuda.synchronize()
P = _lib.PROFILER
P.reset()
P.keep_args = True
P.timing = True
train_step(res[0], model, crit, opt)
torch.cuda.synchronize()
P.timing = False
tot = 0.0
agg = {}
for (name, s, e, kind, work), args in zip(P.records, P.arg_log):
    ms = s.elapsed_time(e)
    tot += ms
    rate = ""
    if kind == "flop" and ms > 0:
        rate = "%7.1f TFLOP/s" % (work / ms / 1e9)
    elif kind == "byte" and ms > 0:
        rate = "%7.1f GB/s" % (work / ms / 1e6)
    key = (name, args)
    a = agg.setdefault(key, [0, 0.0, rate])
    a[0] += 1
    a[1] += ms
print("total of per-call times: %.2f ms" % tot)
for (name, args), (n, ms, rate) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    ints = [a for a in args if abs(a) < 10 ** 7][:12]
    print("%8.3f ms x%-2d %-28s %-16s %s" % (ms, n, name.replace("vocr_", ""), rate, ints))
