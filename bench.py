#!/usr/bin/env python
"""Contract benchmark of the VistaOCR line-recognition hot path on B200.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU implementation of the same step
                                                           # (oracle port of src/train_cnn_lstm.py:131-150 on torch CPU)

Workload (BASELINE.json configs[1]): IAM-style training, line height 60 (rapid-downsample to 30),
batch 64 per GPU, fp32 CNN + 3x512 BiLSTM + CTC, alphabet 96, widths ~ 2*U{150..600}, labels U{20..60}, reference
init U(-0.08,0.08), LSTM dropout 0.5, synthetic images.  A step = forward + CTC + backward + gradient clamp + Adam
(+ NCCL all-reduce for N>1).  `value`: lines/s with the batch already resident in HBM; `e2e`: the same step fed from
pinned host buffers through the public API (H2D copy of the batch and D2H read of the loss inside the timed region).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(input_line_height=60, rds_line_height=30, lstm_input_dim=128, num_lstm_layers=3,
           num_lstm_hidden_units=512, p_lstm_dropout=0.5)
N_SYMBOLS = 96
BATCH = 64
WMIN, WMAX = 300, 1200
LMIN, LMAX = 20, 60
N_BATCHES = 3  # distinct synthetic batches cycled through (= the minimum warm-up, so every shape is seen before timing)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    how="measured")
    return dict(hbm=6650.0, tf=1590.0, tf_sustained=1400.0, how="fallback")


def synth_batches(seed, n, batch=BATCH):
    """SortByWidthCollater contract (reference src/datautils.py:61-176): widths sorted descending, zero right padding,
    values U[0,1), int32 concatenated targets."""
    from vistaocr_b200.ops import out_hw
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        widths = np.sort(rng.integers(WMIN, WMAX + 1, size=batch))[::-1].astype(np.int32).copy()
        x = np.zeros((batch, 1, CFG["input_line_height"], int(widths[0])), np.float32)
        for b in range(batch):
            x[b, :, :, :widths[b]] = rng.random((1, CFG["input_line_height"], widths[b]), dtype=np.float32)
        label_lens = np.zeros(batch, np.int32)
        labels = []
        for b in range(batch):
            t = out_hw(CFG["input_line_height"], int(widths[b]), 1)[1]
            L = int(rng.integers(min(LMIN, t // 2), min(LMAX, t // 2) + 1))
            label_lens[b] = L
            labels.extend(rng.integers(1, N_SYMBOLS, size=L).tolist())
        out.append((torch.from_numpy(x), torch.from_numpy(np.array(labels, np.int32)), torch.from_numpy(widths),
                    torch.from_numpy(label_lens), {}))
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.samples.append([v.strip() for v in o.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if len(s) >= 6 and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) >= 6 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples if len(s) >= 6 for n, v in zip(names, s[2:6]) if v == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU implementation of the same step (oracle port), bounded sample per step
# ------------------------------------------------------------------------------------------------------------------
def cpu_train_lines_per_s(batch, lines, steps, warmup, threads):
    from oracle import model_ref as M
    torch.set_num_threads(threads)
    sd = M.make_state_dict(CFG, N_SYMBOLS, seed=7, lively=False)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running" not in k}
    state = dict(sd)
    state.update(params)
    mom = {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in params.items()}
    x, labels, widths, label_lens, _ = batch
    x, widths, label_lens = x[:lines, :, :, :int(widths[0])], widths[:lines], label_lens[:lines]
    labels = labels[:int(label_lens.sum())]
    g = torch.Generator().manual_seed(7)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        u1, u2 = torch.rand((lines, 64, 2), generator=g), torch.rand((lines, 128, 2), generator=g)
        for p in params.values():
            p.grad = None
        logits, lens = M.forward_ref(state, x, widths.numpy(), CFG, (u1, u2), training=True, bn_updates={})
        loss = M.ctc_sum_ref(logits, labels.numpy(), lens, label_lens.numpy())
        loss.backward()
        with torch.no_grad():
            for k, p in params.items():
                if p.grad is None:
                    continue
                newp, m, v = M.adam_clamp_ref(p, p.grad, mom[k][0], mom[k][1], it + 1)
                p.copy_(newp)
                mom[k] = (m, v)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return lines / float(np.mean(times)), float(np.mean(times))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    lines = 2
    batch = synth_batches(7, 1)[0]
    lps, sec = cpu_train_lines_per_s(batch, lines, args.steps, args.warmup, threads)
    sample = "first %d lines of one synthetic cfg2 batch per step (padded width %d), %d warm-up + %d timed steps" % (
        lines, int(batch[2][0]), args.warmup, args.steps)
    print(json.dumps({
        "impl": "reference", "metric": "train text-lines/sec", "value": lps, "unit": "lines/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": lps, "unit": "lines/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": lps, "unit": "lines/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(n):
    return {"workload": "cfg2: IAM-style training step (fwd + CTC + bwd + clamp + Adam), line height 60 -> rds 30, "
                        "batch 64 per GPU, widths 2*U{150..600}, alphabet 96, labels U{20..60}, D128 / 3x512 BiLSTM, "
                        "dropout 0.5, reference init",
            "global_batch": BATCH * n, "parallelism": "dp%d" % n,
            "l2": "per-step working set (activations ~4 GB) >> 126 MB L2; %d distinct batches cycled" % N_BATCHES}


# ------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no-extras", action="store_true", help="skip the decode / CTC side metrics and the CPU baseline")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from vistaocr_b200 import Alphabet, ClampAdam, CnnOcrModel, CTCLoss, _lib, train_step
    from vistaocr_b200.optim import broadcast_parameters

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pk = peaks()

    torch.manual_seed(7)
    alphabet = Alphabet(["<ctc-blank>"] + ["u%04x" % (0x21 + i) for i in range(N_SYMBOLS - 1)])
    model = CnnOcrModel(alphabet=alphabet, verbose=False, **CFG)
    model.train()
    broadcast_parameters(model)
    torch.manual_seed(7 + rank)  # per-rank fractional-pool samples and dropout masks
    criterion = CTCLoss(host_cost=False)
    optimizer = ClampAdam(model.parameters(), lr=1e-3)
    host = synth_batches(1000 + rank, N_BATCHES)
    pinned = [(b[0].pin_memory(), b[1].pin_memory(), b[2], b[3], b[4]) for b in host]
    resident = [(b[0].to(dev), b[1].to(dev), b[2], b[3], b[4]) for b in host]
    h2d = int(np.mean([b[0].numel() * 4 + b[1].numel() * 4 + b[2].numel() * 4 + b[3].numel() * 4 for b in host]))

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(batches, steps, read_loss):
        sync()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        s.record()
        for i in range(steps):
            loss = train_step(batches[i % len(batches)], model, criterion, optimizer)
            if read_loss:
                float(loss[0].item())  # D2H read of the step's result
        e.record()
        enq = (time.perf_counter() - t0) * 1e3  # host time to enqueue the steps (device still running)
        sync()
        ms = s.elapsed_time(e)
        wall = (time.perf_counter() - t0) * 1e3
        t = torch.tensor([ms, wall, enq], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t[0].item(), t[1].item(), t[2].item()

    # warm-up (allocator, cuFuncSetAttribute, NCCL)
    timed(resident, args.warmup, False)
    _lib.PROFILER.reset()
    with ClockSampler(local_rank) as clk:
        ms, wall, enq = timed(resident, args.steps, False)
    launches = _lib.PROFILER.launches
    ms_e2e, _, _ = timed(pinned, args.steps, True)
    # per-kernel device times: a separate pass with an event pair around every C-ABI call (not part of `value`)
    _lib.PROFILER.reset()
    _lib.PROFILER.timing = True
    timed(resident, args.steps, False)
    _lib.PROFILER.timing = False
    prof = _lib.PROFILER.summary()

    lines = BATCH * world * args.steps
    value = lines / (ms * 1e-3)
    e2e = lines / (ms_e2e * 1e-3)

    # per entry point: device time inside the timed region, algorithmic work, and the roofline that bounds it
    total_kernel_ms = sum(d["ms"] for d in prof.values())
    shares = {k: d["ms"] / total_kernel_ms for k, d in prof.items()}
    NOTES = {
        "vocr_tc_gemm_f16x3": "TMA + tcgen05.mma kind::f16 + TMEM on FP16 pair planes, three compensated products per "
                              "result (1/3 of the f16 rate is the ceiling of this fp32-accurate mode); peak = sustained dense bf16",
        "vocr_tc_conv3x3_fwd_f16": "4-D TMA implicit GEMM + tcgen05 on FP16 pair planes, three products (fwd and data "
                                   "gradient); peak = sustained dense bf16",
        "vocr_tc_conv3x3_wgrad_f16": "4-D TMA implicit GEMM + tcgen05 on FP16 pair planes, chunked TMEM accumulation; "
                                     "peak = sustained dense bf16",
        "vocr_tc_gemm_tf32x3": "TMA + tcgen05.mma kind::tf32 + TMEM, 3xTF32; peak = sustained dense bf16",
        "vocr_tc_conv3x3_fwd": "4-D TMA implicit GEMM + tcgen05 3xTF32 (fwd and data gradient); peak = sustained dense bf16",
        "vocr_tc_conv3x3_wgrad": "4-D TMA implicit GEMM + tcgen05 3xTF32, chunked TMEM accumulation; peak = sustained dense bf16",
        "vocr_bilstm_fwd_f32": "persistent recurrence, W_hh resident in smem as FP16 pairs, h exchanged through L2 flags: "
                               "latency bound (T dependent steps per launch), bytes = T*2*B*5H*4 (SURVEY 8d)",
        "vocr_bilstm_bwd_f32": "persistent recurrence (backward), partial dh reduce-scattered through L2: latency bound, "
                               "bytes = T*2*B*10H*4",
        "vocr_gemm_f32": "fp32 FFMA engine (operands the TMA path cannot address)",
        "vocr_conv3x3_fwd_f32": "fp32 FFMA implicit GEMM (Cin < 32 layers)",
        "vocr_conv3x3_wgrad_f32": "fp32 FFMA implicit GEMM (Cin < 32 layers)",
    }
    # dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full capture of this same cfg2 loop
    # (profiles/r01_f16_kernels_ncu.md); kernels that run with several grids per step have no single figure
    NCU_TRAFFIC = {"vocr_bilstm_bwd_f32": 294.9e6 + 171.4e6, "vocr_bilstm_fwd_f32": 199.4e6 + 247.1e6}

    def roof_of(name):
        d = prof[name]
        if d["kind"] == "flop":
            ach = d["work"] / (d["ms"] * 1e-3) / 1e12
            r = {"kernel": name, "bound": "tensor", "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                 "frac": ach / pk["tf_sustained"]}
        elif d["kind"] == "byte":
            ach = d["work"] / (d["ms"] * 1e-3) / 1e9
            r = {"kernel": name, "bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"]}
        else:
            return None
        r.update({"traffic": NCU_TRAFFIC.get(name), "step_share": shares[name], "avg_launch_ms": d["ms"] / d["calls"],
                  "launches": d["calls"], "note": NOTES.get(name, "") + " (%s peaks)" % pk["how"]})
        return r

    ranked = sorted(prof, key=lambda k: -prof[k]["ms"])
    roof = roof_of(ranked[0])
    rooflines = [r for r in (roof_of(k) for k in ranked[:6]) if r is not None]

    out = {
        "metric": "train text-lines/sec", "value": value, "unit": "lines/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(world),
        "clocks": clk.summary(), "gpu_launches": launches,
        "e2e": {"value": e2e, "unit": "lines/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
        "roofline": roof, "rooflines_top_kernels": rooflines,
        "kernel_time_shares": {k: round(v, 4) for k, v in sorted(shares.items(), key=lambda kv: -kv[1])},
        "host_wall_ms_per_step": wall / args.steps,
        "host_enqueue_ms_per_step": enq / args.steps,
    }

    fp16_mode = None
    if not args.no_extras:  # the same step with fp16 tensor-core operands (one product): cfg3's reduced-precision mode
        import vistaocr_b200
        vistaocr_b200.set_precision("fp16")
        timed(resident, 2, False)
        ms16, _, _ = timed(resident, args.steps, False)
        vistaocr_b200.set_precision("fp32")
        fp16_mode = {"lines_per_s": lines / (ms16 * 1e-3), "ms_per_step": ms16 / args.steps,
                     "what": "same cfg2 step with set_precision('fp16'): GEMM / convolution operands fp16 (hi planes, one "
                             "product), fp32 accumulation, activations, master weights and optimiser; NOT the headline "
                             "value (the metric is quoted in fp32)"}
    if rank == 0 and not args.no_extras:
        out["extra"] = side_metrics(dev, model, alphabet, pk)
        out["extra"]["train_fp16_operands"] = fp16_mode
        if world == 1:
            threads = os.cpu_count() or 1
            lps, sec = cpu_train_lines_per_s(host[0], 2, 2, 1, threads)
            out["cpu_baseline"] = {"value": lps, "unit": "lines/s", "cores": threads, "kind": "port",
                                   "sample": "oracle port of the reference's training step (torch CPU, %d threads) on "
                                             "the first 2 lines of batch 0, 1 warm-up + 2 timed steps" % threads}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def side_metrics(dev, model, alphabet, pk):
    """The other two numbers BASELINE.json's metric names: greedy-decode lines/s (eval forward + decode of 64-line
    cfg1-style batches through the public API) and CTC fwd+bwd GB/s (cfg4 point T=500, A=120, L=50, B=256)."""
    from vistaocr_b200 import CnnOcrModel
    from vistaocr_b200.decoder import greedy_decode_labels
    from vistaocr_b200.warpctc import ctc_costs_and_grads
    res = {}
    flush = torch.zeros(64 << 20, dtype=torch.float32, device=dev)  # 256 MB > L2

    def ev_time(fn, iters, warm=3):
        for _ in range(warm):
            fn()
        ts = []
        for _ in range(iters):
            flush.add_(1.0)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e) * 1e-3)
        return float(np.median(ts))

    g = torch.Generator(device="cuda").manual_seed(7)
    # decode kernel alone at cfg5 scale, and CTC fwd+bwd
    T, B, A = 392, 2048, 120
    x = torch.randn((T, B, A), device=dev, generator=g)
    lens = torch.randint(T // 2, T + 1, (B,), device=dev, generator=g, dtype=torch.int32)
    t = ev_time(lambda: greedy_decode_labels(x, lens, 3 / A), 10)
    res["greedy_decode_kernel"] = {"T": T, "B": B, "A": A, "ms": t * 1e3, "GBs": 4.0 * T * B * A / t / 1e9,
                                   "frac_hbm": 4.0 * T * B * A / t / 1e9 / pk["hbm"], "lines_per_s": B / t}
    del x
    T, B, A, L = 500, 256, 120, 50
    rng = np.random.default_rng(4)
    x = torch.randn((T, B, A), device=dev, generator=g)
    al = torch.from_numpy(np.sort(rng.integers(T // 2, T + 1, size=B))[::-1].astype(np.int32).copy()).to(dev)
    ll = torch.full((B,), L, dtype=torch.int32)
    lab = torch.from_numpy(rng.integers(1, A, size=B * L).astype(np.int32)).to(dev)
    t = ev_time(lambda: ctc_costs_and_grads(x, lab, al, ll), 10)
    res["ctc_fwd_bwd"] = {"T": T, "B": B, "A": A, "L": L, "ms": t * 1e3, "GBs": 8.0 * T * B * A / t / 1e9,
                          "frac_hbm": 8.0 * T * B * A / t / 1e9 / pk["hbm"]}
    del x
    # end-to-end greedy decode: cfg1 (line height 30, 64 lines, widths U{200..800}, alphabet 120) eval forward + decode
    torch.manual_seed(7)
    from vistaocr_b200 import Alphabet
    a1 = Alphabet(["<ctc-blank>"] + ["u%04x" % (0x21 + i) for i in range(119)])
    m1 = CnnOcrModel(alphabet=a1, verbose=False, input_line_height=30, rds_line_height=30, lstm_input_dim=128,
                     num_lstm_layers=3, num_lstm_hidden_units=512, p_lstm_dropout=0.5)
    m1.eval()
    rng = np.random.default_rng(7)
    widths = np.sort(rng.integers(200, 801, size=64))[::-1].astype(np.int32).copy()
    img = np.zeros((64, 1, 30, int(widths[0])), np.float32)
    for b in range(64):
        img[b, :, :, :widths[b]] = rng.random((1, 30, widths[b]), dtype=np.float32)
    img_h = torch.from_numpy(img).pin_memory()
    wt = torch.from_numpy(widths)

    def decode_batch():
        with torch.no_grad():
            logits, lens = m1(img_h.to(dev, non_blocking=True), wt)
            return m1.decode_without_lm(logits, lens, uxxxx=True)

    for _ in range(2):
        decode_batch()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 5
    for _ in range(n):
        decode_batch()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    res["greedy_decode_e2e_cfg1"] = {"lines_per_s": 64 / dt, "ms_per_batch": dt * 1e3,
                                     "what": "H2D + eval forward + greedy decode to strings, 64 lines, host wall clock"}
    del m1
    # cfg5: mixed-width lines (reference bucket mix) in width-bucketed batches of 256 (tools/decode_bench.py; at N GPUs
    # the lines are sharded by bucket with no collective: python -m torch.distributed.run ... tools/decode_bench.py)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools"))
    import decode_bench
    n, dt = decode_bench.run(4096, 256, dev)
    res["greedy_decode_e2e_cfg5"] = {"lines_per_s": n / dt, "lines": n, "batch": 256,
                                     "what": "H2D + eval forward + greedy decode to strings, width-bucketed batches of "
                                             "256 mixed-width lines, host wall clock, 1 GPU"}
    return res


if __name__ == "__main__":
    main()
