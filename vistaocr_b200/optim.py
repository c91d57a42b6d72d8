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


class ClampAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, clamp=5.0,
                 process_group=None, bucket_elems=4 << 20, overlap=True):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, clamp=clamp)
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise ValueError("ClampAdam keeps all parameters in one flat buffer: pass a single parameter group")
        self._params = [p for p in self.param_groups[0]["params"] if p.requires_grad]
        dev = self._params[0].device
        if dev.type != "cuda":
            raise ops._lib.VocrError("ClampAdam needs CUDA parameters: there is no CPU path")
        sizes = [(p.numel() + 3) // 4 * 4 for p in self._params]  # keep every view 16-B aligned
        self._offsets = [0]
        for s in sizes:
            self._offsets.append(self._offsets[-1] + s)
        n = self._offsets[-1]
        self.flat_p = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_v = torch.zeros(n, dtype=torch.float32, device=dev)
        self._gviews = []
        with torch.no_grad():
            for p, o in zip(self._params, self._offsets):
                view = self.flat_p[o:o + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
                gv = self.flat_g[o:o + p.numel()].view_as(p)
                p.grad = gv
                self._gviews.append(gv)
        self._step = 0
        # ---- data parallel plumbing ----
        self._pg = process_group
        self._world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self._handles = []
        self._buckets = []   # (lo, hi) element ranges of flat_g, in the order backward completes them
        self._pending = {}
        if self._world > 1:
            self._make_buckets(bucket_elems)
            if overlap:
                for i, p in enumerate(self._params):
                    p.register_post_accumulate_grad_hook(self._make_hook(i))
            self._overlap = overlap

    # buckets are contiguous ranges of the flat buffer taken from the END (backward reaches the last parameters -
    # prob layer, top LSTM layer - first)
    def _make_buckets(self, bucket_elems):
        self._bucket_of = [0] * len(self._params)
        hi = len(self._params)
        bidx = 0
        while hi > 0:
            lo = hi
            while lo > 0 and self._offsets[hi] - self._offsets[lo] < bucket_elems:
                lo -= 1
            self._buckets.append((self._offsets[lo], self._offsets[hi], hi - lo))
            for i in range(lo, hi):
                self._bucket_of[i] = bidx
            bidx += 1
            hi = lo
        self._reset_pending()

    def _reset_pending(self):
        self._pending = {b: cnt for b, (_, _, cnt) in enumerate(self._buckets)}
        self._launched = set()

    def _launch_bucket(self, b):
        lo, hi, _ = self._buckets[b]
        self._handles.append(dist.all_reduce(self.flat_g[lo:hi], op=dist.ReduceOp.SUM, group=self._pg, async_op=True))
        self._launched.add(b)

    def _make_hook(self, i):
        def hook(p):
            if p.grad is not self._gviews[i]:  # someone replaced .grad (zero_grad(set_to_none=True)): fold it back
                self._gviews[i].copy_(p.grad)
                p.grad = self._gviews[i]
            b = self._bucket_of[i]
            self._pending[b] -= 1
            if self._pending[b] == 0:
                self._launch_bucket(b)
        return hook

    def zero_grad(self, set_to_none=False):
        self.flat_g.zero_()
        for p, gv in zip(self._params, self._gviews):
            p.grad = gv

    @torch.no_grad()
    def step(self, closure=None):
        for p, gv in zip(self._params, self._gviews):
            if p.grad is None:
                gv.zero_()
            elif p.grad is not gv:
                gv.copy_(p.grad)
            p.grad = gv
        if self._world > 1:
            for b in range(len(self._buckets)):
                if b not in self._launched:
                    self._launch_bucket(b)
            for h in self._handles:
                h.wait()
            self._handles = []
            self._reset_pending()
        g = self.param_groups[0]
        self._step += 1
        ops.clamp_adam_step(self.flat_p, self.flat_g, self.flat_m, self.flat_v, self._step, lr=g["lr"],
                            betas=g["betas"], eps=g["eps"], weight_decay=g["weight_decay"], clamp=g["clamp"])

    def gpu_launches_per_step(self):
        return 1


def broadcast_parameters(model, src=0, process_group=None):
    """Replicas start from identical weights (one NCCL broadcast per tensor at load time)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1:
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, src=src, group=process_group)


def train_step(batch, model, criterion, optimizer):
    """The reference's train() (src/train_cnn_lstm.py:131-150) on this package's model / loss / optimizer.
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
