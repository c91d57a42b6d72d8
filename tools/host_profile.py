"""Host-side cost of one cfg2 training step: cProfile over a few steps + implicit-sync warnings (GPU box only)."""
import cProfile, io, os, pstats, sys, time, warnings
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
host = bench.synth_train_batches(bench.TRAIN["train_cfg2"], 1000, 2)
res = [(b[0].to(dev), b[1].to(dev), b[2], b[3], b[4]) for b in host]
for i in range(6):
    train_step(res[i % 2], model, crit, opt)
torch.cuda.synchronize()
torch.cuda.set_sync_debug_mode("warn")
with warnings.catch_warnings(record=True) as w:
    warnings.simplefilter("always")
    train_step(res[0], model, crit, opt)
torch.cuda.set_sync_debug_mode("default")
print("implicit syncs in one step:", len(w))
for x in w[:10]:
    print("  ", str(x.message)[:100], x.filename, x.lineno)
torch.cuda.synchronize()
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
for i in range(5):
    train_step(res[i % 2], model, crit, opt)
pr.disable()
enq = (time.perf_counter() - t0) / 5 * 1e3
torch.cuda.synchronize()
print("host enqueue ms/step (under cProfile): %.2f" % enq)
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(35)
print(s.getvalue()[:6000])
