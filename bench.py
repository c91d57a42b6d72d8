#!/usr/bin/env python
"""Contract benchmark of the VistaOCR line-recognition hot path on B200.

  python bench.py --gpus N --steps K --warmup W [--workload NAME]   # this repo's CUDA path, one process per GPU
  python bench.py --impl reference --gpus N --steps K --warmup W [--workload NAME]
                                         # the reference's own CPU implementation of the same workload (host cores)

Workloads (BASELINE.json configs / SURVEY.md section 8d); each prints ONE JSON line on rank 0 with the contract keys
(metric, value, unit, e2e, roofline, cpu_baseline, clocks, gpu_launches, config ...):
  train_cfg2 (default)  IAM-style training, line height 60 (rapid-downsample to 30), batch 64 per GPU, fp32-contract CNN
                        + 3x512 BiLSTM + CTC, alphabet 96, widths 2*U{150..600}, labels U{20..60}, reference init
                        U(-0.08,0.08), LSTM dropout 0.5.  A step = forward + CTC + backward + gradient clamp + Adam
                        (+ NCCL all-reduce for N > 1).
  train_cfg3            MADCAT-style: line height 120 (two rapid-downsample stages), widths U{400..2000}, alphabet 166,
                        batch 64 per GPU, reduced-precision tensor-core operands (set_precision("fp16")).
  decode_cfg5           greedy-decode throughput: mixed-width lines (reference bucket mix, height 30, alphabet 120) in
                        width-bucketed batches, sharded over the ranks with no collective; 12 500 lines per GPU
                        (100 000 at N = 8).  A step = one batch: H2D + eval forward + greedy decode (+ strings for e2e).
  decode_cfg1           the reference's CPU-runnable case: 64 lines, height 30, widths U{200..800}, alphabet 120.
  ctc_cfg4              standalone CTC forward + gradient over the grid T x A x L at batch 256; value = GB/s of
                        algorithmic bytes (2*T*B*A*4 per point).  A step = one pass over the whole grid.
`value`: the metric with inputs already resident in HBM (device events); `e2e`: the same through the public API fed from
pinned host buffers, H2D of the inputs and D2H of the result inside the timed region.  The training and decode steps
are replayed as CUDA graphs keyed by batch geometry (vistaocr_b200/graphs.py); `--eager` times the eager path instead.
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

TRAIN = {
    "train_cfg2": dict(hp=dict(input_line_height=60, rds_line_height=30, lstm_input_dim=128, num_lstm_layers=3,
                               num_lstm_hidden_units=512, p_lstm_dropout=0.5),
                       n_symbols=96, batch=64, wmin=300, wmax=1200, lmin=20, lmax=60, precision="fp32", dtype="f32",
                       text="cfg2: IAM-style training step (fwd + CTC + bwd + clamp + Adam), line height 60 -> rds 30, "
                            "batch 64 per GPU, widths 2*U{150..600}, alphabet 96, labels U{20..60}, D128 / 3x512 BiLSTM, "
                            "dropout 0.5, reference init; fp32 contract (1e-5) met with 22-bit compensated FP16-pair "
                            "tensor-core operands, fp32 accumulation"),
    "train_cfg3": dict(hp=dict(input_line_height=120, rds_line_height=30, lstm_input_dim=128, num_lstm_layers=3,
                               num_lstm_hidden_units=512, p_lstm_dropout=0.5),
                       n_symbols=166, batch=64, wmin=400, wmax=2000, lmin=20, lmax=60, precision="fp16", dtype="f16",
                       text="cfg3: MADCAT-style training step, line height 120 -> two rapid-downsample stages -> 30, "
                            "batch 64 per GPU, widths U{400..2000}, alphabet 166, labels U{20..60}, D128 / 3x512 BiLSTM, "
                            "dropout 0.5, reference init; reduced precision: fp16 tensor-core operands (per-tensor "
                            "power-of-two scale, 11-bit mantissa vs bf16's 8), fp32 accumulation / activations / master "
                            "weights / optimiser"),
}
N_BATCHES = 3  # distinct synthetic batches cycled through
LSTM_HP = dict(lstm_input_dim=128, num_lstm_layers=3, num_lstm_hidden_units=512, p_lstm_dropout=0.5)
# reference width mix at line height 30 (src/madcat.py:58-66 as quoted in SURVEY.md section 8d): (share, lo, hi)
MIX = [(0.10, 60, 150), (0.10, 150, 200), (0.25, 200, 300), (0.25, 300, 350), (0.20, 350, 450), (0.09, 450, 600),
       (0.01, 600, 1200)]
CFG5_LINES_PER_GPU = 12500
CTC_GRID = [(T, A, L) for T in (100, 250, 500, 1000) for A in (80, 120, 200) for L in (20, 50, 150) if L <= T // 2]
CTC_B = 256


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    how="measured")
    return dict(hbm=6650.0, tf=1590.0, tf_sustained=1400.0, how="fallback")


def alphabet_chars(n):
    return ["<ctc-blank>"] + ["u%04x" % (0x21 + i) for i in range(n - 1)]


def synth_train_batches(cfg, seed, n):
    """SortByWidthCollater contract (reference src/datautils.py:61-176): widths sorted descending, zero right padding,
    values U[0,1), int32 concatenated targets."""
    from vistaocr_b200.ops import out_hw
    hp, batch = cfg["hp"], cfg["batch"]
    h = hp["input_line_height"]
    n_rds = {1: 0, 2: 1, 4: 2, 8: 3}[h // hp["rds_line_height"]]
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        widths = np.sort(rng.integers(cfg["wmin"], cfg["wmax"] + 1, size=batch))[::-1].astype(np.int32).copy()
        x = np.zeros((batch, 1, h, int(widths[0])), np.float32)
        for b in range(batch):
            x[b, :, :, :widths[b]] = rng.random((1, h, widths[b]), dtype=np.float32)
        label_lens = np.zeros(batch, np.int32)
        labels = []
        for b in range(batch):
            t = out_hw(h, int(widths[b]), n_rds)[1]
            L = int(rng.integers(min(cfg["lmin"], t // 2), min(cfg["lmax"], t // 2) + 1))
            label_lens[b] = L
            labels.extend(rng.integers(1, cfg["n_symbols"], size=L).tolist())
        out.append((torch.from_numpy(x), torch.from_numpy(np.array(labels, np.int32)), torch.from_numpy(widths),
                    torch.from_numpy(label_lens), {}))
    return out


def synth_mix_widths(rng, n):
    u = rng.random(n)
    edges = np.cumsum([m[0] for m in MIX])
    which = np.minimum(np.searchsorted(edges, u), len(MIX) - 1)
    lo = np.array([MIX[k][1] for k in which])
    hi = np.array([MIX[k][2] for k in which])
    return (lo + (rng.random(n) * (hi - lo))).astype(np.int32)


def synth_decode_batches(workload, rank, world, batch, max_batches=None):
    """Pinned-host batches of this rank: cfg5 = the reference bucket mix dealt by width bucket over the ranks
    (vistaocr_b200/sharding.py), cfg1 = one fixed 64-line batch per rank."""
    from vistaocr_b200.sharding import shard_batches
    rng = np.random.default_rng(7)
    if workload == "decode_cfg1":
        plans = [np.sort(rng.integers(200, 801, size=64))[::-1].astype(np.int32).copy()]
    else:
        widths = synth_mix_widths(rng, CFG5_LINES_PER_GPU * world)
        plan = shard_batches(widths, 30, batch, world, rank, drop_last=False)
        if max_batches:  # a bounded sample: every k-th batch, so that all width buckets stay represented
            stride = max(1, len(plan) // max_batches)
            plan = plan[::stride][:max_batches]
        plans = [widths[idx] for idx in plan]
    rng = np.random.default_rng(1000 + rank)
    out = []
    for w in plans:
        x = torch.zeros((len(w), 1, 30, int(w[0])), dtype=torch.float32)
        for b, wb in enumerate(w):
            x[b, :, :, :wb] = torch.from_numpy(rng.random((1, 30, int(wb)), dtype=np.float32))
        out.append((x, torch.from_numpy(np.asarray(w, np.int32))))
    return out


def synth_ctc_case(T, A, L, B, seed, dev):
    rng = np.random.default_rng(seed)
    g = torch.Generator(device="cuda").manual_seed(seed) if dev.type == "cuda" else torch.Generator().manual_seed(seed)
    x = torch.randn((T, B, A), device=dev, generator=g)
    al = np.sort(rng.integers(T // 2, T + 1, size=B))[::-1].astype(np.int32).copy()
    ll = np.minimum(L, al // 2).astype(np.int32)
    lab = rng.integers(1, A, size=int(ll.sum())).astype(np.int32)
    rep = rng.random(lab.size) < 0.10  # ~10 % forced repeats
    lab[1:][rep[1:]] = lab[:-1][rep[1:]]
    # feasible by construction: label_len + repeats <= 2 * label_len <= act_len
    return x, torch.from_numpy(lab), torch.from_numpy(al), torch.from_numpy(ll)


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region: NVML (pynvml) every 5 ms when available, else
    nvidia-smi (the recipe's clocks line) as fast as it returns."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index, self.samples, self._stop, self._t, self.how = index, [], threading.Event(), None, "nvidia-smi"

    def _nvml_handle(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].strip().isdigit() else self.index
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
        except Exception:
            return None, None

    def _run(self):
        nv, h = self._nvml_handle()
        if nv is not None:
            self.how = "nvml"
            bits = [getattr(nv, n, 0) for n in ("nvmlClocksEventReasonHwSlowdown", "nvmlClocksEventReasonHwThermalSlowdown",
                                                  "nvmlClocksEventReasonSwThermalSlowdown", "nvmlClocksEventReasonSwPowerCap")]
            if not any(bits):  # older NVML naming
                bits = [getattr(nv, n, 0) for n in ("nvmlClocksThrottleReasonHwSlowdown", "nvmlClocksThrottleReasonHwThermalSlowdown",
                                                      "nvmlClocksThrottleReasonSwThermalSlowdown", "nvmlClocksThrottleReasonSwPowerCap")]
            try:
                mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            except Exception:
                mx = None
            while not self._stop.is_set():
                try:
                    sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                    try:
                        r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                    except Exception:
                        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    self.samples.append([str(sm), str(mx)] + ["Active" if (b and (r & b)) else "Not Active" for b in bits])
                except Exception:
                    pass
                self._stop.wait(0.005)
            return
        while not self._stop.is_set():
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.samples.append([v.strip() for v in o.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(0.05)

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
        reasons = sorted({n for s in self.samples if len(s) >= 6 for n, v in zip(self.NAMES, s[2:6]) if v == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "source": self.how}


class Ctx:
    """Process / device context of one rank."""

    def __init__(self, args):
        import torch.distributed as dist
        self.dist = dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.pk = peaks()
        self.args = args

    def sync(self):
        torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(self, vals):
        t = torch.tensor(vals, dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    def sum_over_ranks(self, vals):
        t = torch.tensor(vals, dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.tolist()

    def timed(self, fn, steps):
        """barrier + synchronize, `steps` calls of fn(i) bracketed by CUDA events, barrier + synchronize; returns
        (device ms, host wall ms, host enqueue ms), each the MAX over ranks."""
        self.sync()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        s.record()
        for i in range(steps):
            fn(i)
        e.record()
        enq = (time.perf_counter() - t0) * 1e3
        self.sync()
        ms = s.elapsed_time(e)
        wall = (time.perf_counter() - t0) * 1e3
        return self.max_over_ranks([ms, wall, enq])

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


NOTES = {
    "vocr_tc_gemm_f16x3": "TMA + tcgen05.mma kind::f16 + TMEM on FP16 pair planes (three compensated products per result "
                          "in the fp32-contract mode: 1/3 of the f16 rate is its ceiling); peak = sustained dense bf16",
    "vocr_tc_conv3x3_fwd_f16": "4-D TMA implicit GEMM + tcgen05 on FP16 pair planes (fwd and data gradient); peak = "
                               "sustained dense bf16",
    "vocr_tc_conv3x3_wgrad_f16": "4-D TMA implicit GEMM + tcgen05 on FP16 pair planes, chunked TMEM accumulation; peak = "
                                 "sustained dense bf16",
    "vocr_bilstm_fwd_f32": "persistent recurrence, W_hh resident in smem as FP16 pairs: latency bound (T dependent steps "
                           "per launch), bytes = T*2*B*5H*4 (SURVEY 8d)",
    "vocr_bilstm_bwd_f32": "persistent recurrence (backward): latency bound, bytes = T*2*B*10H*4",
    "vocr_ctc_loss_f32": "lattice + alpha/beta recursion + gradient; bytes = 2*T*B*A*4 (SURVEY 8d)",
    "vocr_greedy_decode_f32": "arg-max + collapse / compaction; bytes = T*B*A*4",
}
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full captures under profiles/ (cfg2 shapes);
# kernels launched with several shapes per step have no single figure
NCU_TRAFFIC = {"vocr_bilstm_bwd_f32": 294.9e6 + 171.4e6, "vocr_bilstm_fwd_f32": 199.4e6 + 247.1e6,
               # several shapes per step: MEAN over the launches of a cfg2 step (profiles/r02_launches.md)
               "vocr_tc_conv3x3_fwd_f16": 342e6, "vocr_tc_gemm_f16x3": 304e6, "vocr_tc_conv3x3_wgrad_f16": 385e6,
               "vocr_split_f16_f32": 82e6, "vocr_bn_relu_bwd_f32": 1146e6}


def rooflines_from(prof, pk, top=6):
    total = sum(d["ms"] for d in prof.values()) or 1.0
    shares = {k: d["ms"] / total for k, d in prof.items()}

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
    roofs = [r for r in (roof_of(k) for k in ranked) if r is not None]
    return (roofs[0] if roofs else None), roofs[:top], {k: round(v, 4) for k, v in sorted(shares.items(), key=lambda kv: -kv[1])}


def profile_pass(ctx, fn, steps):
    """Separate pass with an event pair around every C-ABI call (not part of `value`)."""
    from vistaocr_b200 import _lib
    _lib.PROFILER.reset()
    _lib.PROFILER.timing = True
    ctx.timed(fn, steps)
    _lib.PROFILER.timing = False
    return _lib.PROFILER.summary()


# ------------------------------------------------------------------------------------------------------------------
# training workloads
# ------------------------------------------------------------------------------------------------------------------
def train_config(name, n, graphs):
    cfg = TRAIN[name]
    return {"workload": cfg["text"], "global_batch": cfg["batch"] * n, "parallelism": "dp%d" % n,
            "l2": "per-step working set (activations, several GB) >> 126 MB L2; %d distinct batches cycled" % N_BATCHES,
            "launch": "CUDA graphs keyed by batch geometry (fwd + CTC + bwd), optimizer step eager" if graphs else
                      "eager (Python -> ctypes per kernel)"}


def run_train(ctx, name):
    import vistaocr_b200
    from vistaocr_b200 import Alphabet, ClampAdam, CnnOcrModel, CTCLoss, GraphedTrainStep, _lib, train_step
    from vistaocr_b200.optim import broadcast_parameters
    args, cfg, dev = ctx.args, TRAIN[name], ctx.dev
    vistaocr_b200.set_precision(cfg["precision"])
    torch.manual_seed(7)
    model = CnnOcrModel(alphabet=Alphabet(alphabet_chars(cfg["n_symbols"])), verbose=False, **cfg["hp"])
    model.train()
    broadcast_parameters(model)
    torch.manual_seed(7 + ctx.rank)  # per-rank fractional-pool samples and dropout masks
    criterion = CTCLoss(host_cost=False)
    optimizer = ClampAdam(model.parameters(), lr=1e-3)
    host = synth_train_batches(cfg, 1000 + ctx.rank, N_BATCHES)
    pinned = [(b[0].pin_memory(), b[1].pin_memory(), b[2], b[3], b[4]) for b in host]
    resident = [(b[0].to(dev), b[1].to(dev), b[2], b[3], b[4]) for b in host]
    h2d = int(np.mean([b[0].numel() * 4 + b[1].numel() * 4 + b[2].numel() * 4 + b[3].numel() * 4 for b in host]))
    graphed = GraphedTrainStep(model, criterion, optimizer, capture_after=1, max_graphs=2 * N_BATCHES)
    use_graphs = not args.eager

    def step_fn(batches, read_loss, eager=False):
        def fn(i):
            b = batches[i % len(batches)]
            loss = train_step(b, model, criterion, optimizer) if (eager or not use_graphs) else graphed(b)
            if read_loss:
                float(loss[0].item())  # D2H read of the step's result
        return fn

    # warm-up: every geometry is seen once eagerly (library / allocator / NCCL warm-up) and captured on its second visit
    warm = max(args.warmup, 2 * N_BATCHES if use_graphs else 3)
    ctx.timed(step_fn(resident, False), warm)
    _lib.PROFILER.reset()
    with ClockSampler(ctx.local_rank) as clk:
        ms, wall, enq = ctx.timed(step_fn(resident, False), args.steps)
    launches_api = _lib.PROFILER.launches
    ms_e2e, _, _ = ctx.timed(step_fn(pinned, True), args.steps)
    # per-kernel device times and the launch count of one step: an eager pass (the graph replays exactly these launches)
    prof = profile_pass(ctx, step_fn(resident, False, eager=True), max(3, min(args.steps, 6)))
    _lib.PROFILER.reset()
    ms_eager, _, enq_eager = ctx.timed(step_fn(resident, False, eager=True), max(3, min(args.steps, 6)))
    eager_steps = max(3, min(args.steps, 6))
    launches_per_step = _lib.PROFILER.launches / eager_steps
    roof, roofs, shares = rooflines_from(prof, ctx.pk)

    lines = cfg["batch"] * ctx.world * args.steps
    out = {
        "metric": "train text-lines/sec", "value": lines / (ms * 1e-3), "unit": "lines/s", "n_gpus": ctx.world,
        "steps": args.steps, "warmup": warm, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": cfg["dtype"], "data": "synthetic", "config": train_config(name, ctx.world, use_graphs),
        "clocks": clk.summary(),
        "gpu_launches": int(round(launches_per_step * args.steps)) if use_graphs else launches_api,
        "gpu_launches_note": "kernels of this repo's library executed in the timed region (%d per step, counted on the "
                             "eager path; the graph replays the same launches)" % round(launches_per_step),
        "e2e": {"value": lines / (ms_e2e * 1e-3), "unit": "lines/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
        "roofline": roof, "rooflines_top_kernels": roofs, "kernel_time_shares": shares,
        "host_wall_ms_per_step": wall / args.steps, "host_enqueue_ms_per_step": enq / args.steps,
        "eager": {"ms_per_step": ms_eager / eager_steps, "host_enqueue_ms_per_step": enq_eager / eager_steps,
                  "lines_per_s": cfg["batch"] * ctx.world * eager_steps / (ms_eager * 1e-3)},
        "graphs": {"captures": graphed.graphs.captures, "replays": graphed.graphs.replays,
                   "eager_calls": graphed.graphs.eager_calls},
    }
    if not args.no_extras and name == "train_cfg2":
        # the same step with fp16 tensor-core operands (one product): cfg3's reduced-precision mode on cfg2 shapes
        vistaocr_b200.set_precision("fp16")
        ctx.timed(step_fn(resident, False), 2 * N_BATCHES)
        ms16, _, _ = ctx.timed(step_fn(resident, False), args.steps)
        vistaocr_b200.set_precision("fp32")
        out["extra"] = {"train_fp16_operands": {
            "lines_per_s": lines / (ms16 * 1e-3), "ms_per_step": ms16 / args.steps,
            "what": "same cfg2 step with set_precision('fp16'); NOT the headline value (cfg2 is quoted in fp32)"}}
        if ctx.rank == 0:
            out["extra"].update(side_metrics(ctx))
    if ctx.rank == 0 and ctx.world == 1 and not args.no_extras:
        out["cpu_baseline"] = cpu_train_baseline(name, host[0], cfg["batch"], 1, 0)
    vistaocr_b200.set_precision("fp32")
    return out


def cpu_train_baseline(name, batch, lines, steps, warmup):
    """The reference's train() on the host cores for `lines` lines of one synthetic batch: the reference's own CnnOcrModel
    (oracle/ref_runner.py; kind "reference") or, where the staged reference is absent, the oracle port (kind "port")."""
    from oracle import ref_runner as R
    cfg = TRAIN[name]
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    x, labels, widths, label_lens, _ = batch
    x, widths, label_lens = x[:lines, :, :, :int(widths[0])].contiguous(), widths[:lines], label_lens[:lines]
    labels = labels[:int(label_lens.sum())]
    ns = R.load()
    times = []
    if ns is not None:
        model = R.make_model(ns, cfg["hp"], alphabet_chars(cfg["n_symbols"]))
        model.train()
        opt = torch.optim.Adam(model.parameters(), lr=1e-3)
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            R.train_step(model, opt, (x, labels, widths, label_lens))
            if it >= warmup:
                times.append(time.perf_counter() - t0)
        kind = "reference"
        what = "the reference's own CnnOcrModel + train() sequence (warp-ctc replaced by torch CPU ctc_loss)"
    else:
        from oracle import model_ref as M
        sd = M.make_state_dict(cfg["hp"], cfg["n_symbols"], seed=7, lively=False)
        params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running" not in k}
        state = dict(sd)
        state.update(params)
        mom = {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in params.items()}
        g = torch.Generator().manual_seed(7)
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            u1, u2 = torch.rand((lines, 64, 2), generator=g), torch.rand((lines, 128, 2), generator=g)
            for p in params.values():
                p.grad = None
            logits, lens = M.forward_ref(state, x, widths.numpy(), cfg["hp"], (u1, u2), training=True, bn_updates={})
            M.ctc_sum_ref(logits, labels.numpy(), lens, label_lens.numpy()).backward()
            with torch.no_grad():
                for k, p in params.items():
                    if p.grad is not None:
                        newp, m, v = M.adam_clamp_ref(p, p.grad, mom[k][0], mom[k][1], it + 1)
                        p.copy_(newp)
                        mom[k] = (m, v)
            if it >= warmup:
                times.append(time.perf_counter() - t0)
        kind, what = "port", "oracle port of the reference's training step"
    sec = float(np.mean(times))
    return {"value": lines / sec, "unit": "lines/s", "cores": threads, "kind": kind, "seconds_per_step": sec,
            "sample": "%s on torch CPU, %d threads: %d lines of one synthetic %s batch (padded width %d), %d warm-up + %d "
                      "timed steps" % (what, threads, lines, name, int(widths[0]), warmup, steps)}


# ------------------------------------------------------------------------------------------------------------------
# decode workloads
# ------------------------------------------------------------------------------------------------------------------
def decode_model(n_symbols=120):
    from vistaocr_b200 import Alphabet, CnnOcrModel
    torch.manual_seed(7)
    m = CnnOcrModel(alphabet=Alphabet(alphabet_chars(n_symbols)), verbose=False, input_line_height=30,
                    rds_line_height=30, **LSTM_HP)
    m.eval()
    return m


def run_decode(ctx, name):
    from vistaocr_b200 import GraphedDecoder, _lib
    from vistaocr_b200.decoder import greedy_decode_labels
    from vistaocr_b200.optim import broadcast_parameters
    args, dev = ctx.args, ctx.dev
    batch = 64 if name == "decode_cfg1" else args.batch
    model = decode_model()
    broadcast_parameters(model)
    host = synth_decode_batches(name, ctx.rank, ctx.world, batch, max_batches=args.steps if name == "decode_cfg5" else None)
    steps = args.steps if name == "decode_cfg1" else len(host)
    pinned = [(x.pin_memory(), w) for x, w in host]
    resident = [(x.to(dev), w) for x, w in host]
    # cfg5 geometries do not repeat: its GraphedDecoder never captures and runs its (fused) eager path
    use_graphs = (not args.eager) and name == "decode_cfg1"
    graphed = GraphedDecoder(model, capture_after=1 if use_graphs else 1 << 30, max_graphs=4)
    uncaptured = GraphedDecoder(model, capture_after=1 << 30)  # the same call sequence, launched kernel by kernel
    thresh = 3 * 1 / len(model.alphabet)

    def device_fn(i):  # resident inputs, result stays on the device
        x, w = resident[i % len(resident)]
        if not args.eager:
            (graphed if use_graphs else uncaptured).labels(x, w, allow_eager=True)
            return
        with torch.no_grad():
            logits, lens = model(x, w)
            greedy_decode_labels(logits, lens, thresh)

    def e2e_fn(i):  # pinned host batch in, strings out
        x, w = pinned[i % len(pinned)]
        if not args.eager:
            graphed(x, w, uxxxx=True)
        else:
            with torch.no_grad():
                logits, lens = model(x.to(dev, non_blocking=True), w)
                model.decode_without_lm(logits, lens, uxxxx=True)

    warm = max(args.warmup, 3)
    # cfg5: every batch has its own geometry - one untimed pass over all of them (allocator, first-use costs), then W more
    ctx.timed(device_fn, warm if name == "decode_cfg1" else len(host) + warm)
    _lib.PROFILER.reset()
    with ClockSampler(ctx.local_rank) as clk:
        ms, wall, enq = ctx.timed(device_fn, steps)
    launches = _lib.PROFILER.launches
    ctx.timed(e2e_fn, min(steps, 2))
    ms_e2e, _, _ = ctx.timed(e2e_fn, steps)
    use_graphs_saved, use_graphs = use_graphs, False
    prof = profile_pass(ctx, device_fn, min(steps, 8))
    _lib.PROFILER.reset()
    ctx.timed(device_fn, min(steps, 8))
    launches_eager = _lib.PROFILER.launches / min(steps, 8)
    use_graphs = use_graphs_saved
    roof, roofs, shares = rooflines_from(prof, ctx.pk)
    my_lines = sum(len(host[i % len(host)][1]) for i in range(steps))
    lines = ctx.sum_over_ranks([float(my_lines)])[0]
    h2d = int(np.mean([x.numel() * 4 + w.numel() * 4 for x, w in host]))
    d2h = int(np.mean([2 * 4 * x.shape[0] * max(1, int(w[0]) * 49 // 100) for x, w in host]))  # labels [B,T] + counts
    text = ("cfg1: greedy CTC decode of 64 synthetic lines, line height 30, widths U{200..800}, alphabet 120, D128 / 3x512 "
            "BiLSTM, reference init; a step = eval forward + greedy decode of the batch" if name == "decode_cfg1" else
            "cfg5: %d mixed-width lines per GPU (reference bucket mix, height 30, alphabet 120) in width-bucketed batches "
            "of %d, sharded by bucket over the ranks, no collective; a step = one batch (eval forward + greedy decode)%s"
            % (CFG5_LINES_PER_GPU, batch, "; bounded sample of %d batches per rank" % steps
               if steps * batch < CFG5_LINES_PER_GPU else ""))
    out = {
        "metric": "greedy-decode lines/sec", "value": lines / (ms * 1e-3), "unit": "lines/s", "n_gpus": ctx.world,
        "steps": steps, "warmup": warm, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": text, "global_batch": batch * ctx.world, "parallelism": "replicas x%d (lines sharded)" % ctx.world,
                   "l2": "activations of a batch >> 126 MB L2; distinct batches every step" if name == "decode_cfg5" else
                         "activations of the batch (~1 GB) >> 126 MB L2",
                   "launch": "CUDA graph per batch geometry" if use_graphs else "eager"},
        "clocks": clk.summary(), "gpu_launches": launches if not use_graphs else int(round(launches_eager * steps)),
        "e2e": {"value": lines / (ms_e2e * 1e-3), "unit": "lines/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "roofline": roof, "rooflines_top_kernels": roofs, "kernel_time_shares": shares,
        "host_wall_ms_per_step": wall / steps, "host_enqueue_ms_per_step": enq / steps, "lines": int(lines),
    }
    if ctx.rank == 0 and ctx.world == 1 and not args.no_extras:
        out["cpu_baseline"] = cpu_decode_baseline(host[0], 64)
    return out


def cpu_decode_baseline(batch, lines):
    from oracle import ref_runner as R
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    x, w = batch
    x, w = x[:lines, :, :, :int(w[0])].contiguous(), w[:lines]
    ns = R.load()
    if ns is not None:
        model = R.make_model(ns, dict(input_line_height=30, rds_line_height=30, **LSTM_HP), alphabet_chars(120))
        model.eval()
        dec = ns.ArgmaxDecoder(model.alphabet)
        fn = lambda: R.decode_batch(model, dec, x, w)
        kind, what = "reference", "the reference's own CnnOcrModel.eval() forward + ArgmaxDecoder.decode"
    else:
        from oracle import model_ref as M
        from oracle.decode_ref import decode_loop
        hp = dict(input_line_height=30, rds_line_height=30, **LSTM_HP)
        sd = M.make_state_dict(hp, 120, seed=7, lively=False)
        chars = alphabet_chars(120)
        g = torch.Generator().manual_seed(7)

        def fn():
            with torch.no_grad():
                u = (torch.rand((len(w), 64, 2), generator=g), torch.rand((len(w), 128, 2), generator=g))
                logits, lens = M.forward_ref(sd, x, w.numpy(), hp, u, training=False)
            return decode_loop(logits.numpy(), lens.numpy(), dict(enumerate(chars)), uxxxx=True)
        kind, what = "port", "oracle port of the reference's eval forward + greedy decode"
    fn()
    times = []
    for _ in range(2):
        t0 = time.perf_counter()
        fn()
        times.append(time.perf_counter() - t0)
    sec = float(np.mean(times))
    return {"value": len(w) / sec, "unit": "lines/s", "cores": threads, "kind": kind, "seconds_per_step": sec,
            "sample": "%s on torch CPU, %d threads: one batch of %d lines (padded width %d), 1 warm-up + 2 timed"
                      % (what, threads, len(w), int(w[0]))}


# ------------------------------------------------------------------------------------------------------------------
# CTC workload
# ------------------------------------------------------------------------------------------------------------------
def run_ctc(ctx, name):
    from vistaocr_b200 import CTCLoss, _lib
    from vistaocr_b200.warpctc import ctc_costs_and_grads
    args, dev = ctx.args, ctx.dev
    cases = []
    for i, (T, A, L) in enumerate(CTC_GRID):
        x, lab, al, ll = synth_ctc_case(T, A, L, CTC_B, 40 + i, dev)
        cases.append(dict(T=T, A=A, L=L, x=x, lab=lab.to(dev), al=al.to(dev), ll=ll, lab_h=lab.pin_memory(),
                          al_h=al, x_h=None, bytes=8.0 * T * CTC_B * A))
    total_bytes = sum(c["bytes"] for c in cases)  # 2.2 GB of logits + gradients per pass: >> L2

    def device_fn(i):
        for c in cases:
            ctc_costs_and_grads(c["x"], c["lab"], c["al"], c["ll"])

    warm = max(args.warmup, 3)
    ctx.timed(device_fn, warm)
    _lib.PROFILER.reset()
    with ClockSampler(ctx.local_rank) as clk:
        ms, wall, enq = ctx.timed(device_fn, args.steps)
    launches = _lib.PROFILER.launches
    # per grid point (events around each call, L2 flushed by the 100+ MB the other points move in between)
    per = []
    for c in cases:
        ts = []
        for _ in range(3):
            device_fn(0)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            ctc_costs_and_grads(c["x"], c["lab"], c["al"], c["ll"])
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        t = float(np.median(ts)) * 1e-3
        per.append({"T": c["T"], "A": c["A"], "L": c["L"], "us": t * 1e6, "GBs": c["bytes"] / t / 1e9,
                    "frac_hbm": c["bytes"] / t / 1e9 / ctx.pk["hbm"]})
    # e2e: the public CTCLoss module from pinned host activations, loss read back (a subset of the grid: the host
    # copies of all 33 activation tensors would be 1.1 GB of pinned memory)
    crit = CTCLoss(host_cost=True)
    sub = cases[::4]
    for c in sub:
        c["x_h"] = c["x"].cpu().pin_memory()
    sub_bytes = sum(c["bytes"] for c in sub)

    def e2e_fn(i):
        for c in sub:
            xd = c["x_h"].to(dev, non_blocking=True).requires_grad_(True)
            loss = crit(xd, c["lab_h"], c["al_h"], c["ll"])
            loss.backward()
            xd.grad.cpu()  # the gradient is the result the caller consumes (warp-ctc hands it to autograd)

    ctx.timed(e2e_fn, 1)
    ms_e2e, _, _ = ctx.timed(e2e_fn, max(1, args.steps // 2))
    n_e2e = max(1, args.steps // 2)
    prof = profile_pass(ctx, device_fn, 2)
    roof, roofs, shares = rooflines_from(prof, ctx.pk)
    gbs = total_bytes * args.steps * ctx.world / (ms * 1e-3) / 1e9
    out = {
        "metric": "CTC fwd+bwd GB/s", "value": gbs, "unit": "GB/s", "n_gpus": ctx.world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "cfg4: standalone CTC loss forward + gradient, grid T in {100,250,500,1000} x A in "
                               "{80,120,200} x L in {20,50,150} (L <= T/2: %d points), batch 256, act_lens U{T/2..T}, ~10 %% "
                               "forced label repeats; a step = one pass over the grid; bytes = 2*T*B*A*4 per point"
                               % len(CTC_GRID),
                   "global_batch": CTC_B * ctx.world, "parallelism": "replicas x%d" % ctx.world,
                   "l2": "one pass moves %.1f GB of activations + gradients through HBM, >> 126 MB L2" % (total_bytes / 1e9)},
        "clocks": clk.summary(), "gpu_launches": launches,
        "e2e": {"value": sub_bytes * n_e2e * ctx.world / (ms_e2e * 1e-3) / 1e9, "unit": "GB/s",
                "h2d_bytes_per_step": int(sub_bytes / 2), "d2h_bytes_per_step": int(sub_bytes / 2),
                "note": "every 4th grid point through CTCLoss from pinned host activations, gradient copied back: PCIe bound"},
        "roofline": roof, "rooflines_top_kernels": roofs, "kernel_time_shares": shares, "grid": per,
        "host_enqueue_ms_per_step": enq / args.steps,
    }
    if ctx.rank == 0 and ctx.world == 1 and not args.no_extras:
        out["cpu_baseline"] = cpu_ctc_baseline(2)
    return out


def cpu_ctc_baseline(passes, points=None):
    """torch.nn.functional.ctc_loss forward + backward on the host cores (warp-ctc itself is not installable; this is
    the CPU path the reference's container offers for the same call) over a bounded sample of the grid."""
    import torch.nn.functional as F
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    grid = points or CTC_GRID[::4]
    cases = [synth_ctc_case(T, A, L, CTC_B, 40 + i, torch.device("cpu")) + (8.0 * T * CTC_B * A,)
             for i, (T, A, L) in enumerate(grid)]
    times = []
    for it in range(1 + passes):
        t0 = time.perf_counter()
        for x, lab, al, ll, _ in cases:
            x = x.clone().requires_grad_(True)
            F.ctc_loss(x.log_softmax(2), lab.long(), al.long(), ll.long(), blank=0, reduction="sum",
                       zero_infinity=True).backward()
        if it >= 1:
            times.append(time.perf_counter() - t0)
    sec = float(np.mean(times))
    byt = sum(c[4] for c in cases)
    return {"value": byt / sec / 1e9, "unit": "GB/s", "cores": threads, "kind": "port", "seconds_per_step": sec,
            "sample": "torch CPU ctc_loss(log_softmax) fwd+bwd, %d threads, %d of the %d grid points (every 4th), 1 warm-up "
                      "+ %d timed passes" % (threads, len(cases), len(CTC_GRID), passes)}


# ------------------------------------------------------------------------------------------------------------------
# side metrics of the default line (the other two numbers of BASELINE.json's metric, at one point each)
# ------------------------------------------------------------------------------------------------------------------
def side_metrics(ctx):
    from vistaocr_b200 import GraphedDecoder
    from vistaocr_b200.decoder import greedy_decode_labels
    from vistaocr_b200.warpctc import ctc_costs_and_grads
    dev, pk = ctx.dev, ctx.pk
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
    T, B, A = 392, 2048, 120
    x = torch.randn((T, B, A), device=dev, generator=g)
    lens = torch.randint(T // 2, T + 1, (B,), device=dev, generator=g, dtype=torch.int32)
    t = ev_time(lambda: greedy_decode_labels(x, lens, 3 / A), 10)
    res["greedy_decode_kernel"] = {"T": T, "B": B, "A": A, "ms": t * 1e3, "GBs": 4.0 * T * B * A / t / 1e9,
                                   "frac_hbm": 4.0 * T * B * A / t / 1e9 / pk["hbm"], "lines_per_s": B / t}
    del x
    T, A, L = 500, 120, 50
    x, lab, al, ll = synth_ctc_case(T, A, L, CTC_B, 4, dev)
    lab, al = lab.to(dev), al.to(dev)
    t = ev_time(lambda: ctc_costs_and_grads(x, lab, al, ll), 10)
    res["ctc_fwd_bwd"] = {"T": T, "B": CTC_B, "A": A, "L": L, "ms": t * 1e3, "GBs": 8.0 * T * CTC_B * A / t / 1e9,
                          "frac_hbm": 8.0 * T * CTC_B * A / t / 1e9 / pk["hbm"]}
    del x
    # cfg1 decode end to end (pinned host batch in, strings out), graph replay and eager
    m1 = decode_model()
    (xb, wb), = synth_decode_batches("decode_cfg1", 0, 1, 64)
    xb = xb.pin_memory()
    gd = GraphedDecoder(m1, capture_after=1)

    def wall(fn, n=8):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n

    dt = wall(lambda: gd(xb, wb, uxxxx=True))
    res["greedy_decode_e2e_cfg1"] = {"lines_per_s": 64 / dt, "ms_per_batch": dt * 1e3,
                                     "what": "H2D + eval forward + greedy decode to strings, 64 lines, CUDA-graph replay, "
                                             "host wall clock"}
    gd_eager = GraphedDecoder(m1, capture_after=1 << 30)  # the same call sequence launched kernel by kernel
    dt = wall(lambda: gd_eager(xb, wb, uxxxx=True))
    res["greedy_decode_e2e_cfg1_eager"] = {"lines_per_s": 64 / dt, "ms_per_batch": dt * 1e3}
    return res


# ------------------------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU implementation of the same workload, bounded sample per step
# ------------------------------------------------------------------------------------------------------------------
def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    name, n = args.workload, args.steps + args.warmup
    if name in TRAIN:
        lines = TRAIN[name]["batch"] if n <= 3 else (16 if n <= 8 else 8)  # keeps the whole run within a few minutes
        batch = synth_train_batches(TRAIN[name], 1000, 1)[0]
        cb = cpu_train_baseline(name, batch, lines, args.steps, args.warmup)
        config = train_config(name, args.gpus, False)
        config["launch"] = "CPU"
        config["reference_sample"] = "%d lines per step (of the workload's %d-line batch)" % (lines, TRAIN[name]["batch"])
        metric, unit, dtype = "train text-lines/sec", "lines/s", TRAIN[name]["dtype"]
    elif name in ("decode_cfg5", "decode_cfg1"):
        host = synth_decode_batches(name, 0, 1, 64, max_batches=1)
        cb = cpu_decode_baseline(host[0], 64)
        config = {"workload": name + ": the same synthetic lines through the reference's eval forward + ArgmaxDecoder",
                  "reference_sample": "one 64-line batch per step"}
        metric, unit, dtype = "greedy-decode lines/sec", "lines/s", "f32"
    else:
        cb = cpu_ctc_baseline(max(1, min(args.steps, 3)))
        config = {"workload": "cfg4: CTC fwd+bwd on the CPU", "reference_sample": cb["sample"]}
        metric, unit, dtype = "CTC fwd+bwd GB/s", "GB/s", "f32"
    print(json.dumps({
        "impl": "reference", "metric": metric, "value": cb["value"], "unit": unit, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["seconds_per_step"] * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic", "config": config, "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="train_cfg2",
                    choices=["train_cfg2", "train_cfg3", "decode_cfg5", "decode_cfg1", "ctc_cfg4"])
    ap.add_argument("--batch", type=int, default=448,
                    help="decode_cfg5: lines per batch (a multiple of 224 = 7 clusters x 32 samples: the BiLSTM forward "
                         "kernel keeps 7 clusters of 16 CTAs resident, so 448 lines are exactly four rounds per direction)")
    ap.add_argument("--eager", action="store_true", help="time the eager launch path instead of CUDA-graph replay")
    ap.add_argument("--no-extras", action="store_true", help="skip the side metrics and the in-line CPU baseline")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    ctx = Ctx(args)
    if args.workload in TRAIN:
        out = run_train(ctx, args.workload)
    elif args.workload == "ctc_cfg4":
        out = run_ctc(ctx, args.workload)
    else:
        out = run_decode(ctx, args.workload)
    if ctx.rank == 0:
        print(json.dumps(out))
    ctx.close()


if __name__ == "__main__":
    main()
