"""CUDA-graph replay of the two per-batch call sequences of the hot path (B200: "capture launch-bound inner loops in
CUDA graphs"): the training step (reference src/train_cnn_lstm.py:131-150) and eval forward + greedy decode
(src/decode_testset.py:92-101,166).  A step is ~400 kernel launches issued through Python -> ctypes; replaying them as
one graph removes that host work from the critical path.

What a graph bakes in is the batch GEOMETRY - (B, C, H, padded width, max frame count, longest labelling, mode) - not
the data: pixels, per-line lengths and labels are copied into static device buffers before every replay, the kernels
read lengths from the device, fractional-pool samples come from torch's graph-safe generator and the dropout mask from a
device-resident Philox state that the graph itself advances.  Graphs are cached per geometry (LRU); a geometry runs
eagerly the first `capture_after` times it is seen (that also warms the library up), so streams of never-repeating
widths simply stay on the eager path.  The optimizer step (host-side step count) and, for N > 1, the gradient
all-reduce run eagerly right after the replay.
"""
from collections import OrderedDict

import torch

from . import ops
from .decoder import _canon_map, greedy_decode_labels, collapse_labels, labels_to_strings
from .optim import ClampAdam, train_step
from .warpctc import _CTC, count_infeasible


def _lens_for(model, widths):
    h = model.input_line_height
    return [ops.out_hw(h, int(w), model.num_rds_layers)[1] for w in widths]


class _Entry:
    pass


class _GraphCache:
    def __init__(self, capture_after, max_graphs):
        self.capture_after, self.max_graphs = capture_after, max_graphs
        self.cache, self.seen, self.pool = OrderedDict(), {}, None
        self.captures = self.replays = self.eager_calls = 0

    def lookup(self, key):
        e = self.cache.get(key)
        if e is not None:
            self.cache.move_to_end(key)
            return e, True
        n = self.seen[key] = self.seen.get(key, 0) + 1
        return None, n > self.capture_after

    def store(self, key, e):
        self.cache[key] = e
        self.captures += 1
        while len(self.cache) > self.max_graphs:
            self.cache.popitem(last=False)

    def capture(self, fn):
        g = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, pool=self.pool):
            out = fn()
        if self.pool is None:
            self.pool = g.pool()  # later graphs share the memory pool: only one of them replays at a time
        return g, out


class GraphedTrainStep:
    """`step = GraphedTrainStep(model, criterion, optimizer); loss = step(batch)` - same contract as
    vistaocr_b200.train_step(batch, model, criterion, optimizer)."""

    def __init__(self, model, criterion, optimizer, capture_after=1, max_graphs=8):
        self.model, self.criterion, self.optimizer = model, criterion, optimizer
        self.graphs = _GraphCache(capture_after, max_graphs)

    def _eligible(self):
        c = self.criterion
        return isinstance(self.optimizer, ClampAdam) and not c.host_cost and not c.length_average and \
            getattr(self.model, "_dropout_masks", None) is None and self.model.training

    def __call__(self, batch):
        x, target, widths, target_widths, _ = batch
        if not self._eligible():
            self.graphs.eager_calls += 1
            return train_step(batch, self.model, self.criterion, self.optimizer)
        model = self.model
        wl = widths.tolist() if torch.is_tensor(widths) else list(widths)
        lens = _lens_for(model, wl)
        if any(lens[i] < lens[i + 1] for i in range(len(lens) - 1)):
            raise RuntimeError("`actual_minibatch_widths` must be sorted in decreasing order")
        ll = torch.as_tensor(target_widths).to(torch.int32)
        max_l = int(ll.max().item()) if ll.numel() else 0
        key = (tuple(x.shape), max(lens), max_l, ops.get_precision(), self.criterion.size_average)
        e, want_capture = self.graphs.lookup(key)
        if e is None and not want_capture:
            self.graphs.eager_calls += 1
            return train_step(batch, model, self.criterion, self.optimizer)
        if e is None:
            e = self._capture(key, x, wl, max_l)
        # ---- stage this batch into the graph's static buffers (stream ordered, nothing synchronises) ----
        n_lab = int(target.numel())
        if n_lab > e.labels.numel() or x.shape[3] < max(wl):
            raise ops._lib.VocrError("batch does not fit the captured geometry")
        e.x.copy_(x, non_blocking=True)
        e.lens.copy_(torch.tensor(lens, dtype=torch.int32))
        if n_lab:
            e.labels[:n_lab].copy_(torch.as_tensor(target).to(torch.int32))
        e.label_lens.copy_(ll)
        self.criterion.num_infeasible = count_infeasible(target, torch.tensor(lens, dtype=torch.int32), ll)
        e.graph.replay()
        self.graphs.replays += 1
        self.optimizer.step()  # eager: all-reduce (N > 1) + fused clamp / Adam with the host-side step count
        return e.loss.clone()

    def _capture(self, key, x, widths, max_l):
        model, opt = self.model, self.optimizer
        dev = next(model.parameters()).device
        B = x.shape[0]
        e = _Entry()
        e.x = torch.zeros(tuple(x.shape), dtype=torch.float32, device=dev)
        e.lens = torch.zeros((B,), dtype=torch.int32, device=dev)
        e.labels = torch.zeros((max(1, B * max_l),), dtype=torch.int32, device=dev)
        e.label_lens = torch.zeros((B,), dtype=torch.int32, device=dev)
        scale = 1.0 / B if self.criterion.size_average else 1.0
        # everything that is created lazily must exist before the capture (a captured initialisation would only run
        # at replay time)
        ops.const_scalar(dev, 1.0)
        if model.p_lstm_dropout > 0:
            ops.const_scalar(dev, 1.0 / (1.0 - model.p_lstm_dropout))
            model._rng_state(dev)
        widths_cpu = torch.tensor(widths, dtype=torch.int32)

        def body():
            opt.zero_grad()
            logits, _ = model(e.x, widths_cpu)
            loss = _CTC.apply(logits, e.labels, e.lens, e.label_lens, scale, False, max_l)
            loss.backward()
            return loss.detach()

        model._lens_dev_override = e.lens
        opt.reducer.capturing = True
        try:
            e.graph, e.loss = self.graphs.capture(body)
        finally:
            model._lens_dev_override = None
            opt.reducer.capturing = False
        self.graphs.store(key, e)
        return e


class GraphedDecoder:
    """`decode = GraphedDecoder(model); strings = decode(x, widths, uxxxx=False)`: model.eval() forward + greedy decode
    (= `model.decode_without_lm(*model(x, widths))`), replayed as one graph per batch geometry."""

    def __init__(self, model, capture_after=1, max_graphs=8):
        self.model = model
        self.graphs = _GraphCache(capture_after, max_graphs)
        self._canon = "unset"

    def _canon_dev(self, dev):
        if self._canon == "unset":
            c = _canon_map(self.model.alphabet)
            self._canon = None if c is None else torch.from_numpy(c).to(dev)
        return self._canon

    def _forward_labels(self, x_dev, widths_cpu, lens_dev):
        """Eval forward + greedy decode on the device: (labels, counts).  The frame labels come straight from the
        prob-layer GEMM's epilogue when the shapes allow it (the logits are then never written)."""
        model = self.model
        thresh = 3 * 1 / len(model.alphabet)
        canon = self._canon_dev(x_dev.device)
        with torch.no_grad():
            fd = model._fused_decode = {"thresh": thresh}
            try:
                logits, _ = model(x_dev, widths_cpu)
            finally:
                model._fused_decode = None
            if logits is None:  # arg-max done in the GEMM epilogue: only the collapse is left
                return collapse_labels(fd["path"], lens_dev, canon)
            labels, counts, _ = greedy_decode_labels(logits, lens_dev, thresh, canon)
            return labels, counts

    def labels(self, x, widths, allow_eager=False):
        """Device result of one batch: (labels[B,T] int32, counts[B] int32).  Replayed geometries return the graph's
        static buffers (valid until the next call with the same geometry); a geometry that is not captured (yet) runs
        eagerly when allow_eager is set, else None is returned."""
        model = self.model
        if model.training:
            return None
        wl = widths.tolist() if torch.is_tensor(widths) else list(widths)
        lens = _lens_for(model, wl)
        if any(lens[i] < lens[i + 1] for i in range(len(lens) - 1)):
            raise RuntimeError("`actual_minibatch_widths` must be sorted in decreasing order")
        key = (tuple(x.shape), max(lens), ops.get_precision())
        e, want_capture = self.graphs.lookup(key)
        if e is None and not want_capture:
            if not allow_eager:
                return None
            self.graphs.eager_calls += 1
            dev = next(model.parameters()).device
            lens_dev = torch.tensor(lens, dtype=torch.int32).to(dev, non_blocking=True)
            model._lens_dev_override = lens_dev
            try:
                return self._forward_labels(x.to(dev, non_blocking=True), torch.tensor(wl, dtype=torch.int32), lens_dev)
            finally:
                model._lens_dev_override = None
        if e is None:
            e = self._capture(key, x, wl)
        e.x.copy_(x, non_blocking=True)
        e.lens.copy_(torch.tensor(lens, dtype=torch.int32))
        e.graph.replay()
        self.graphs.replays += 1
        return e.labels, e.counts

    def __call__(self, x, widths, uxxxx=False):
        if self.model.training:
            self.graphs.eager_calls += 1
            with torch.no_grad():
                logits, lens = self.model(x.cuda(non_blocking=True), widths)
                return self.model.decode_without_lm(logits, lens, uxxxx=uxxxx)
        out = self.labels(x, widths, allow_eager=True)
        return labels_to_strings(out[0], out[1], self.model.alphabet, uxxxx)

    def _capture(self, key, x, widths):
        model = self.model
        dev = next(model.parameters()).device
        self._canon_dev(dev)
        e = _Entry()
        e.x = torch.zeros(tuple(x.shape), dtype=torch.float32, device=dev)
        e.lens = torch.zeros((x.shape[0],), dtype=torch.int32, device=dev)
        ops.const_scalar(dev, 1.0)
        widths_cpu = torch.tensor(widths, dtype=torch.int32)

        def body():
            return self._forward_labels(e.x, widths_cpu, e.lens)

        model._lens_dev_override = e.lens
        try:
            e.graph, (e.labels, e.counts) = self.graphs.capture(body)
        finally:
            model._lens_dev_override = None
        self.graphs.store(key, e)
        return e
