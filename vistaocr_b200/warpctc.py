"""Drop-in for `warpctc_pytorch.CTCLoss` (reference call sites src/train_cnn_lstm.py:12,52,138,358).

`CTCLoss()(acts[T,B,A] raw activations on the GPU, labels int32 1-D CPU, act_lens int32[B] CPU,
label_lens int32[B] CPU) -> Tensor[1]` = summed negative log-likelihood; the gradient w.r.t. `acts`
(softmax - occupancy, zero beyond act_lens) is computed in the forward pass by csrc/ctc.cu and handed to
autograd in backward, scaled by grad_output.  warp-ctc keywords `size_average` / `length_average` are honoured.
"""
import torch
import torch.nn as nn

from . import _lib


def ctc_costs_and_grads(acts, labels, act_lens, label_lens, want_grads=True, max_label_len=None):
    """Device part: returns (costs[B] fp32 CUDA, grads[T,B,A] fp32 CUDA or None).  No host sync (device-resident
    label lengths need `max_label_len`, an upper bound of the longest labelling, to size the workspace without one)."""
    _lib.require_cuda(acts, "acts", torch.float32)
    T, B, A = acts.shape
    dev = acts.device
    label_lens_h = torch.as_tensor(label_lens).to(torch.int32).cpu() if not torch.is_tensor(label_lens) or \
        not label_lens.is_cuda else None
    if max_label_len is not None and label_lens_h is None:
        max_l = int(max_label_len)
        label_lens_d = label_lens.to(torch.int32)
    elif label_lens_h is not None:
        max_l = int(label_lens_h.max().item()) if label_lens_h.numel() else 0
        label_lens_d = label_lens_h.to(dev, non_blocking=True)
    else:  # already on the device: one sync to size the workspace
        max_l = int(label_lens.max().item()) if label_lens.numel() else 0
        label_lens_d = label_lens.to(torch.int32)
    labels_d = torch.as_tensor(labels).to(device=dev, dtype=torch.int32, non_blocking=True).contiguous()
    act_lens_d = torch.as_tensor(act_lens).to(device=dev, dtype=torch.int32, non_blocking=True).contiguous()
    if label_lens_d.numel() != B or act_lens_d.numel() != B:
        raise _lib.VocrError("act_lens / label_lens must have one entry per batch element")
    l = _lib.lib()
    ws_bytes = l.vocr_ctc_workspace_size(T, B, A, max_l)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    costs = torch.empty((B,), dtype=torch.float32, device=dev)
    grads = torch.empty_like(acts) if want_grads else None
    st = l.vocr_ctc_loss_f32(_lib.ptr(acts), _lib.ptr(grads), _lib.ptr(labels_d) if labels_d.numel() else None,
                             _lib.ptr(label_lens_d), _lib.ptr(act_lens_d), T, B, A, max_l, _lib.ptr(costs),
                             _lib.ptr(ws), ws_bytes, _lib.stream())
    _lib.check(st, "vocr_ctc_loss_f32")
    return costs, grads


class _CTC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, acts, labels, act_lens, label_lens, scale, host_cost, max_label_len=None):
        acts_c = acts.detach().contiguous()
        costs, grads = ctc_costs_and_grads(acts_c, labels, act_lens, label_lens, want_grads=acts.requires_grad,
                                           max_label_len=max_label_len)
        if grads is not None and scale != 1.0:
            grads.mul_(scale)
        ctx.grads = grads
        total = costs.sum().reshape(1) * scale
        return total.cpu() if host_cost else total

    @staticmethod
    def backward(ctx, grad_output):
        g = ctx.grads
        ctx.grads = None
        go = grad_output.to(g.device, non_blocking=True).reshape(())
        return g * go, None, None, None, None, None, None


def count_infeasible(labels, act_lens, label_lens):
    """Utterances no alignment exists for (act_len < label_len + number of adjacent repeated labels); host arithmetic
    on the CPU tensors the reference passes (datautils.py:159-164).  Returns None when the inputs live on the device."""
    if any(torch.is_tensor(t) and t.is_cuda for t in (labels, act_lens, label_lens)):
        return None
    import numpy as np
    lab = torch.as_tensor(labels).numpy()
    al, ll = torch.as_tensor(act_lens).numpy().astype(np.int64), torch.as_tensor(label_lens).numpy().astype(np.int64)
    ends = np.cumsum(ll)
    starts = ends - ll
    eq = np.zeros(len(lab) + 1, np.int64)
    if len(lab) > 1:
        eq[1:len(lab)] = lab[1:] == lab[:-1]
    eq[starts] = 0  # the first label of an utterance repeats nothing
    cs = np.concatenate([[0], np.cumsum(eq[:len(lab)])])
    rep = cs[ends] - cs[starts]
    return int((al < ll + rep).sum())


class CTCLoss(nn.Module):
    """host_cost=True returns the cost as a CPU FloatTensor[1] exactly like warp-ctc (one D2H sync);
    host_cost=False keeps it on the device so the training step stays asynchronous.
    infeasible: what an utterance with no valid alignment contributes - "zero" (cost 0 / gradient 0, so one bad line
    cannot poison a batch) or "inf" (the cost becomes +inf, gradient still 0: bad label / length data shows up in the
    logged loss).  Either way `num_infeasible` holds the count for the last call (None for device-resident lengths)."""

    def __init__(self, size_average=False, length_average=False, host_cost=True, infeasible="zero"):
        super().__init__()
        if infeasible not in ("zero", "inf"):
            raise ValueError("infeasible must be 'zero' or 'inf'")
        self.size_average = size_average
        self.length_average = length_average
        self.host_cost = host_cost
        self.infeasible = infeasible
        self.num_infeasible = 0

    def forward(self, acts, labels, act_lens, label_lens):
        assert labels.dim() == 1 and act_lens.dim() == 1 and label_lens.dim() == 1
        scale = 1.0
        if self.length_average:
            scale = 1.0 / float(torch.as_tensor(act_lens).sum().item())
        elif self.size_average:
            scale = 1.0 / acts.size(1)
        self.num_infeasible = count_infeasible(labels, act_lens, label_lens)
        loss = _CTC.apply(acts, labels, act_lens, label_lens, scale, self.host_cost)
        if self.infeasible == "inf" and self.num_infeasible:
            loss = loss + float("inf")
        return loss
