"""Training-step tail of the hot path (reference src/train_cnn_lstm.py:139-150,363): gradient all-reduce across data
parallel ranks, element-wise clamp to [-5,5] and the Adam update - here ONE fused kernel over a flat parameter
buffer instead of ~120 per-tensor launches, with the NCCL all-reduce of the flat gradient buffer launched bucket by
bucket from autograd hooks so it overlaps the rest of backward.

`ClampAdam` keeps torch.optim.Optimizer's surface (param_groups / zero_grad / step / state_dict) so the reference's
`optimizer = torch.optim.Adam(model.parameters(), lr=..., weight_decay=...)` line is the only one that changes; the
reference's separate `param.grad.data.clamp_(-5, 5)` loop stays valid (the fused clamp is idempotent).
Sum (not mean) across ranks: the reference's loss is a batch SUM, so N ranks x 64 lines == one batch of 64N lines.
"""
import torch
import torch.distributed as dist

from . import ops


class FlatGradReducer:
    """One flat fp32 gradient buffer for a list of parameters (each `p.grad` is a view into it) plus the data-parallel
    exchange: contiguous buckets of the buffer are all-reduced (SUM) asynchronously as soon as autograd has produced
    every gradient of the bucket, last parameters first (prob layer, top LSTM layer ... CNN), so the collective
    overlaps the remaining backward.  Device-agnostic (NCCL on GPUs; the gloo tests drive it on CPU)."""

    def __init__(self, params, process_group=None, bucket_elems=4 << 20, overlap=True):
        self.params = [p for p in params if p.requires_grad]
        dev = self.params[0].device
        sizes = [(p.numel() + 3) // 4 * 4 for p in self.params]  # keep every view 16-B aligned
        self.offsets = [0]
        for s in sizes:
            self.offsets.append(self.offsets[-1] + s)
        self.flat_g = torch.zeros(self.offsets[-1], dtype=torch.float32, device=dev)
        self.gviews = []
        for p, o in zip(self.params, self.offsets):
            gv = self.flat_g[o:o + p.numel()].view_as(p)
            p.grad = gv
            self.gviews.append(gv)
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.handles = []
        self.capturing = False
        self.buckets = []  # (lo, hi, n_params) element ranges of flat_g, in the order backward completes them
        self.collectives = 0
        if self.world > 1:
            self._make_buckets(bucket_elems)
            if overlap:
                for i, p in enumerate(self.params):
                    p.register_post_accumulate_grad_hook(self._make_hook(i))

    def _make_buckets(self, bucket_elems):
        self.bucket_of = [0] * len(self.params)
        hi = len(self.params)
        while hi > 0:
            lo = hi - 1
            while lo > 0 and self.offsets[hi] - self.offsets[lo] < bucket_elems:
                lo -= 1
            for i in range(lo, hi):
                self.bucket_of[i] = len(self.buckets)
            self.buckets.append((self.offsets[lo], self.offsets[hi], hi - lo))
            hi = lo
        self._reset()

    def _reset(self):
        self.pending = [cnt for (_, _, cnt) in self.buckets]
        self.launched = set()

    def _launch(self, b):
        lo, hi, _ = self.buckets[b]
        self.handles.append(dist.all_reduce(self.flat_g[lo:hi], op=dist.ReduceOp.SUM, group=self.pg, async_op=True))
        self.launched.add(b)
        self.collectives += 1

    def _make_hook(self, i):
        def hook(p):
            if self.capturing:  # CUDA-graph capture (graphs.py): the exchange runs after the replay, in finish()
                return
            if p.grad is not self.gviews[i]:  # someone replaced .grad (zero_grad(set_to_none=True)): fold it back
                self.gviews[i].copy_(p.grad)
                p.grad = self.gviews[i]
            b = self.bucket_of[i]
            if b in self.launched:
                # a second backward before step() (gradient accumulation, retain_graph) would add local gradients into
                # a buffer that is already reduced / in flight and the ranks would silently diverge
                raise RuntimeError("FlatGradReducer: gradient of parameter %d arrived after its bucket was all-reduced; "
                                   "overlap=True supports exactly one backward per optimizer.step() - construct "
                                   "ClampAdam(..., overlap=False) to accumulate over several backward passes" % i)
            self.pending[b] -= 1
            if self.pending[b] == 0:
                self._launch(b)
        return hook

    def zero(self):
        self.flat_g.zero_()
        for p, gv in zip(self.params, self.gviews):
            p.grad = gv

    def finish(self):
        """Fold stray .grad tensors back into the flat buffer, launch whatever has not been launched, wait."""
        for p, gv in zip(self.params, self.gviews):
            if p.grad is None:
                gv.zero_()
            elif p.grad is not gv:
                gv.copy_(p.grad)
            p.grad = gv
        if self.world > 1:
            for b in range(len(self.buckets)):
                if b not in self.launched:
                    self._launch(b)
            for h in self.handles:
                h.wait()
            self.handles = []
            self._reset()


class ClampAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, clamp=5.0,
                 process_group=None, bucket_elems=4 << 20, overlap=True):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, clamp=clamp)
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise ValueError("ClampAdam keeps all parameters in one flat buffer: pass a single parameter group")
        plist = [p for p in self.param_groups[0]["params"] if p.requires_grad]
        if plist[0].device.type != "cuda":
            raise ops._lib.VocrError("ClampAdam needs CUDA parameters: there is no CPU path")
        self.reducer = FlatGradReducer(plist, process_group, bucket_elems, overlap)
        n = self.reducer.offsets[-1]
        dev = plist[0].device
        self.flat_p = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_v = torch.zeros(n, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, o in zip(plist, self.reducer.offsets):
                view = self.flat_p[o:o + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
        self._step = 0

    @property
    def flat_g(self):
        return self.reducer.flat_g

    # ---- checkpointing: the reference snapshots `optimizer.state_dict()` (train_cnn_lstm.py:427-438) ---------------
    def _views(self, flat):
        plist = self.reducer.params
        return [flat[o:o + p.numel()].view_as(p) for p, o in zip(plist, self.reducer.offsets)]

    def state_dict(self):
        """torch.optim.Adam's format (per-parameter `step`, `exp_avg`, `exp_avg_sq`), so a snapshot written here loads
        into torch.optim.Adam and vice versa; the tensors are copies of the flat moment buffers."""
        self.state.clear()
        if self._step > 0:
            for p, m, v in zip(self.reducer.params, self._views(self.flat_m), self._views(self.flat_v)):
                self.state[p] = {"step": torch.tensor(float(self._step)), "exp_avg": m.clone(), "exp_avg_sq": v.clone()}
        sd = super().state_dict()
        self.state.clear()
        return sd

    @torch.no_grad()
    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)  # validates groups / sizes, casts the tensors to the parameters' device
        steps = set()
        self.flat_m.zero_()
        self.flat_v.zero_()
        for p, m, v in zip(self.reducer.params, self._views(self.flat_m), self._views(self.flat_v)):
            st = self.state.get(p)
            if st:
                m.copy_(st["exp_avg"])
                v.copy_(st["exp_avg_sq"])
                steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise ValueError("ClampAdam keeps one step count for the whole flat buffer; the snapshot has %s" % sorted(steps))
        self._step = steps.pop() if steps else 0
        self.state.clear()

    def zero_grad(self, set_to_none=False):
        self.reducer.zero()

    @torch.no_grad()
    def step(self, closure=None):
        self.reducer.finish()
        g = self.param_groups[0]
        self._step += 1
        ops.clamp_adam_step(self.flat_p, self.reducer.flat_g, self.flat_m, self.flat_v, self._step, lr=g["lr"],
                            betas=g["betas"], eps=g["eps"], weight_decay=g["weight_decay"], clamp=g["clamp"])


def broadcast_parameters(model, src=0, process_group=None):
    """Replicas start from identical weights (one NCCL broadcast per tensor at load time)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1:
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, src=src, group=process_group)


def train_step(batch, model, criterion, optimizer):
    """The reference's train() (src/train_cnn_lstm.py:131-150) on this package's model / loss / optimizer.
    Data-parallel runs must give every rank the same number of steps (sharding.shard_batches(drop_last=True)): the
    gradient all-reduce is a collective, a rank with one batch fewer would leave the others waiting in it.
    Returns the loss tensor (index [0] like the reference's `loss.data[0]`) without forcing a host sync."""
    input_tensor, target, input_widths, target_widths, metadata = batch
    input_tensor = input_tensor.cuda(non_blocking=True)
    optimizer.zero_grad()
    model_output, model_output_actual_lengths = model(input_tensor, input_widths)
    loss = criterion(model_output, target, model_output_actual_lengths, target_widths)
    loss.backward()
    if not isinstance(optimizer, ClampAdam):
        for param in model.parameters():
            if param.grad is not None:
                param.grad.data.clamp_(min=-5, max=5)
    optimizer.step()
    return loss.data
