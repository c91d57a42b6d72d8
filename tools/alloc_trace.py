"""Caching-allocator behaviour across cfg2 training steps (GPU box only): cudaMalloc/cudaFree counts per step."""
import os, sys, time
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
nb = int(os.environ.get("NB", 3))
host = bench.synth_train_batches(bench.TRAIN["train_cfg2"], 1000, nb)
res = [(b[0].to(dev), b[1].to(dev), b[2], b[3], b[4]) for b in host]
prev = torch.cuda.memory_stats()
for i in range(int(os.environ.get("STEPS", 15))):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    train_step(res[i % nb], model, crit, opt)
    enq = time.perf_counter() - t0
    torch.cuda.synchronize()
    tot = time.perf_counter() - t0
    st = torch.cuda.memory_stats()
    print("step %2d batch %d W=%d: enqueue %.1f ms total %.1f ms  cudaMalloc +%d cudaFree +%d retries +%d  reserved %.1f GB peak-alloc %.1f GB" % (
        i, i % nb, res[i % nb][0].shape[-1], enq * 1e3, tot * 1e3,
        st["num_device_alloc"] - prev["num_device_alloc"], st["num_device_free"] - prev["num_device_free"],
        st["num_alloc_retries"] - prev["num_alloc_retries"], st["reserved_bytes.all.current"] / 2**30,
        st["allocated_bytes.all.peak"] / 2**30))
    prev = st
