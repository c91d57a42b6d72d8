"""Oracle (test infrastructure): the reference's CnnOcrModel forward / training step restated with plain PyTorch
CPU ops, functional over a reference-format state_dict.  Each function cites the reference lines it follows.

The reference delegates this path to torch.nn modules (src/models/cnnlstm.py:114-154), so the restatement is in
torch fp32/fp64 on CPU too.  Pinned against the reference itself, imported from /root/reference by
tests/golden/make_golden.py (fixtures tests/golden/model_*.npz) and tests/test_oracle_vs_reference.py.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

CONV_IDX = (0, 3, 7, 10, 14, 17, 20)  # positions of the Conv2d modules inside reference `cnn` (cnnlstm.py:124-134)
BN_IDX = (1, 4, 8, 11, 15, 18, 21)
POOL_AFTER = (1, 3)                   # FractionalMaxPool2d follows conv blocks 2 and 4 (modules 6 and 13)
CONV_CH = (64, 64, 128, 128, 256, 256, 256)


def num_rds_layers(input_line_height, rds_line_height):
    # cnnlstm.py:96-112
    n, lh = 0, input_line_height
    while lh > rds_line_height:
        n += 1
        lh //= 2
    return n


def out_hw(h, w, n_rds):
    """cnn_input_size_to_output_size (cnnlstm.py:211-260): conv3x3/pad1 keeps size, MaxPool2d(2,2) floors the
    half, FractionalMaxPool2d floors x*0.5 / x*0.7 in float64 (so 350 -> 244, not 245)."""
    for _ in range(n_rds):
        h, w = math.floor((h - 2) / 2 + 1), math.floor((w - 2) / 2 + 1)
    for _ in range(2):
        h, w = math.floor(h * 0.5), math.floor(w * 0.7)
    return h, w


def fmp_starts(u, in_size, out_size, dtype=np.float32):
    """ATen fractional_max_pool2d interval generation for pool size 2, computed in the input's scalar type:
    alpha = (in-2)/(out-1); start_i = int((i+u)*alpha) - int(u*alpha) for i < out-1; start_{out-1} = in-2."""
    u = dtype(u)
    alpha = dtype(in_size - 2) / dtype(out_size - 1)
    seq = np.empty(out_size, np.int64)
    for i in range(out_size - 1):
        seq[i] = int(dtype(dtype(i) + u) * alpha) - int(dtype(u * alpha))
    seq[out_size - 1] = in_size - 2
    return seq


def fmp_ref(x, samples, interval_dtype=np.float32):
    """FractionalMaxPool2d(2, output_ratio=(0.5,0.7)) with explicit per-(n,c) samples [N,C,2]
    (samples[...,0] drives W, [...,1] drives H), via the formula above.  x: [N,C,H,W].  The window positions are
    computed in `interval_dtype` - float32 is what the (fp32) reference model does; keeping it for a float64 `x` gives
    the float64 evaluation of the SAME function (F.fractional_max_pool2d on float64 input would move some windows)."""
    N, C, H, W = x.shape
    Ho, Wo = int(H * 0.5), int(W * 0.7)
    s = samples.detach().cpu().numpy().astype(np.float32)
    planes = []
    for n in range(N):
        for c in range(C):
            ws = torch.from_numpy(fmp_starts(s[n, c, 0], W, Wo, interval_dtype)).to(x.device)
            hs = torch.from_numpy(fmp_starts(s[n, c, 1], H, Ho, interval_dtype)).to(x.device)
            p = x[n, c]
            rows = torch.maximum(p[hs], p[hs + 1])
            planes.append(torch.maximum(rows[:, ws], rows[:, ws + 1]))
    return torch.stack(planes).view(N, C, Ho, Wo)


def cnn_ref(sd, x, n_rds, pool_samples, training, bn_updates=None, prefix="cnn."):
    """rapid_ds + cnn (cnnlstm.py:114-134, 270-271).  pool_samples = (u1[N,64,2], u2[N,128,2]).
    In training mode BatchNorm uses batch statistics (over the zero-padded region too) and, if `bn_updates` is a
    dict, the updated running stats are returned in it (momentum 0.1, unbiased variance)."""
    for i in range(n_rds):
        x = F.conv2d(x, sd["rapid_ds.%02d-conv.weight" % i], sd["rapid_ds.%02d-conv.bias" % i], padding=1)
        x = F.max_pool2d(F.relu(x), 2, stride=2)
    for k, (ci, bi) in enumerate(zip(CONV_IDX, BN_IDX)):
        x = F.conv2d(x, sd["%s%d.weight" % (prefix, ci)], sd["%s%d.bias" % (prefix, ci)], padding=1)
        rm, rv = sd["%s%d.running_mean" % (prefix, bi)], sd["%s%d.running_var" % (prefix, bi)]
        if training and bn_updates is not None:
            rm, rv = rm.clone(), rv.clone()
            bn_updates["%s%d.running_mean" % (prefix, bi)] = rm
            bn_updates["%s%d.running_var" % (prefix, bi)] = rv
        x = F.batch_norm(x, None if (training and bn_updates is None) else rm,
                         None if (training and bn_updates is None) else rv,
                         sd["%s%d.weight" % (prefix, bi)], sd["%s%d.bias" % (prefix, bi)], training, 0.1, 1e-5)
        x = F.relu(x)
        if k in POOL_AFTER:
            u = pool_samples[POOL_AFTER.index(k)]
            if x.dtype == torch.float32:
                x = F.fractional_max_pool2d(x, 2, output_ratio=(0.5, 0.7), _random_samples=u.to(x.dtype))
            else:  # float64 twin of the fp32 model: same (float32-computed) windows
                x = fmp_ref(x, u)
    return x


def lstm_cell_loop(x, lens, w_ih, w_hh, b_ih, b_hh, reverse):
    """One direction of one LSTM layer, explicit time loop with packed-sequence semantics (cnnlstm.py:288-290):
    gate order i,f,g,o; h0=c0=0; the reverse direction starts at each sample's own last valid frame; outputs at
    t >= len are zero.  x: [T,B,D] -> [T,B,H]."""
    T, B, _ = x.shape
    H = w_hh.shape[1]
    h = x.new_zeros((B, H))
    c = x.new_zeros((B, H))
    out = [None] * T
    lens_t = torch.as_tensor(lens).to(x.device)
    for k in range(T):
        t = T - 1 - k if reverse else k
        gates = x[t] @ w_ih.t() + b_ih + h @ w_hh.t() + b_hh
        i, f, g, o = gates.chunk(4, dim=1)
        c_new = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h_new = torch.sigmoid(o) * torch.tanh(c_new)
        m = (t < lens_t).to(x.dtype)[:, None]
        c = m * c_new + (1 - m) * c
        h = m * h_new + (1 - m) * h
        out[t] = m * h_new
    return torch.stack(out, 0)


def bilstm_ref(sd, x, lens, num_layers, dropout_masks=None):
    """Stacked BiLSTM (cnnlstm.py:148-149); dropout_masks[l] (already scaled by 1/(1-p)) multiplies the output of
    layer l < num_layers-1, as nn.LSTM's inter-layer dropout does in training."""
    for l in range(num_layers):
        outs = []
        for suffix, rev in (("", False), ("_reverse", True)):
            outs.append(lstm_cell_loop(x, lens, sd["lstm.weight_ih_l%d%s" % (l, suffix)],
                                       sd["lstm.weight_hh_l%d%s" % (l, suffix)],
                                       sd["lstm.bias_ih_l%d%s" % (l, suffix)],
                                       sd["lstm.bias_hh_l%d%s" % (l, suffix)], rev))
        x = torch.cat(outs, 2)
        if dropout_masks is not None and l < num_layers - 1:
            x = x * dropout_masks[l]
    return x


def bilstm_nn(sd, x, lens, num_layers, hidden, dtype):
    """Same computation through torch.nn.LSTM on a packed sequence - literally what the reference runs."""
    lstm = torch.nn.LSTM(x.shape[2], hidden, num_layers=num_layers, bidirectional=True).to(dtype)
    weights = {k[len("lstm."):]: v for k, v in sd.items() if k.startswith("lstm.")}
    packed = torch.nn.utils.rnn.pack_padded_sequence(x, list(lens))
    out, _ = torch.func.functional_call(lstm, weights, (packed,))  # keeps autograd attached to `sd`'s tensors
    out, _ = torch.nn.utils.rnn.pad_packed_sequence(out)
    return out


def forward_ref(sd, x, widths, hp, pool_samples, training=False, bn_updates=None, dropout_masks=None,
                use_nn_lstm=True):
    """CnnOcrModel.forward (cnnlstm.py:268-296).  sd: reference-format state_dict (cnn.N.* keys), x [B,C,H,W],
    widths [B] sorted descending.  Returns (logits [T',B,A], lens int32 [B])."""
    n_rds = num_rds_layers(hp["input_line_height"], hp["rds_line_height"])
    feat = cnn_ref(sd, x, n_rds, pool_samples, training, bn_updates)
    b, c, h, w = feat.shape
    seq = feat.permute(3, 0, 1, 2).contiguous().view(-1, c * h)
    seq = F.relu(F.linear(seq, sd["bridge_layer.0.weight"], sd["bridge_layer.0.bias"])).view(w, b, -1)
    lens = [out_hw(hp["input_line_height"], int(wd), n_rds)[1] for wd in widths]
    L, Hd = hp["num_lstm_layers"], hp["num_lstm_hidden_units"]
    tmax = max(lens)
    if use_nn_lstm and dropout_masks is None:
        out = bilstm_nn(sd, seq, lens, L, Hd, seq.dtype)
    else:
        out = bilstm_ref(sd, seq, lens, L, dropout_masks)[:tmax]
    logits = F.linear(out.reshape(-1, out.shape[2]), sd["prob_layer.0.weight"], sd["prob_layer.0.bias"])
    return logits.view(out.shape[0], b, -1), torch.tensor(lens, dtype=torch.int32)


def ctc_sum_ref(logits, labels, act_lens, label_lens):
    """warp-ctc semantics on CPU (train_cnn_lstm.py:138): summed cost, softmax inside, infeasible -> 0."""
    return F.ctc_loss(logits.log_softmax(2), torch.as_tensor(labels, dtype=torch.long),
                      torch.as_tensor(act_lens, dtype=torch.long), torch.as_tensor(label_lens, dtype=torch.long),
                      blank=0, reduction="sum", zero_infinity=True)


def adam_clamp_ref(p, g, m, v, step, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8, wd=0.0, clamp=5.0):
    """Element-wise grad clamp to [-5,5] (train_cnn_lstm.py:143-145) followed by torch.optim.Adam's update
    (:363; L2 added to the gradient, bias-corrected, eps outside the sqrt of the corrected second moment)."""
    g = g.clamp(-clamp, clamp)
    if wd != 0.0:
        g = g + wd * p
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    p = p - (lr / bc1) * m / ((v.sqrt() / math.sqrt(bc2)) + eps)
    return p, m, v


def make_state_dict(hp, n_symbols, seed, lively=True, dtype=torch.float32):
    """Deterministic reference-format state_dict from numpy's PCG64 stream (stable across platforms), so fixtures
    need not store 17 M weights.  lively=False reproduces the reference init U(-0.08,0.08) on every parameter
    (cnnlstm.py:158-159); lively=True uses fan-in scaled weights and BN gamma near 1 so argmax paths move."""
    rng = np.random.default_rng(seed)
    n_rds = num_rds_layers(hp["input_line_height"], hp["rds_line_height"])
    sd = {}

    def u(shape, bound):
        return torch.from_numpy(rng.uniform(-bound, bound, size=shape).astype(np.float32)).to(dtype)

    cin = hp.get("num_in_channels", 1)
    for i in range(n_rds):
        b = math.sqrt(3.0 / (cin * 9)) if lively else 0.08
        sd["rapid_ds.%02d-conv.weight" % i] = u((16, cin, 3, 3), b)
        sd["rapid_ds.%02d-conv.bias" % i] = u((16,), 0.1 if lively else 0.08)
        cin = 16
    for ci, bi, co in zip(CONV_IDX, BN_IDX, CONV_CH):
        b = math.sqrt(6.0 / (cin * 9)) if lively else 0.08
        sd["cnn.%d.weight" % ci] = u((co, cin, 3, 3), b)
        sd["cnn.%d.bias" % ci] = u((co,), 0.1 if lively else 0.08)
        sd["cnn.%d.weight" % bi] = (1.0 + u((co,), 0.3)) if lively else u((co,), 0.08)
        sd["cnn.%d.bias" % bi] = u((co,), 0.3 if lively else 0.08)
        sd["cnn.%d.running_mean" % bi] = u((co,), 0.2) if lively else torch.zeros(co, dtype=dtype)
        sd["cnn.%d.running_var" % bi] = (1.0 + u((co,), 0.5)) if lively else torch.ones(co, dtype=dtype)
        sd["cnn.%d.num_batches_tracked" % bi] = torch.tensor(0, dtype=torch.long)
        cin = co
    h_out = out_hw(hp["input_line_height"], 20, n_rds)[0]
    feat = 256 * h_out
    D, Hd, L = hp["lstm_input_dim"], hp["num_lstm_hidden_units"], hp["num_lstm_layers"]
    sd["bridge_layer.0.weight"] = u((D, feat), math.sqrt(6.0 / feat) if lively else 0.08)
    sd["bridge_layer.0.bias"] = u((D,), 0.1 if lively else 0.08)
    for l in range(L):
        din = D if l == 0 else 2 * Hd
        for suffix in ("", "_reverse"):
            k = (1.5 / math.sqrt(Hd)) if lively else 0.08
            sd["lstm.weight_ih_l%d%s" % (l, suffix)] = u((4 * Hd, din), (1.5 / math.sqrt(din)) if lively else 0.08)
            sd["lstm.weight_hh_l%d%s" % (l, suffix)] = u((4 * Hd, Hd), k)
            sd["lstm.bias_ih_l%d%s" % (l, suffix)] = u((4 * Hd,), 0.2 if lively else 0.08)
            sd["lstm.bias_hh_l%d%s" % (l, suffix)] = u((4 * Hd,), 0.2 if lively else 0.08)
    sd["prob_layer.0.weight"] = u((n_symbols, 2 * Hd), (3.0 / math.sqrt(2 * Hd)) if lively else 0.08)
    sd["prob_layer.0.bias"] = u((n_symbols,), 0.5 if lively else 0.08)
    return sd


def synth_batch(rng, B, H, wmin, wmax, n_symbols, lmin=0, lmax=0, n_rds=0):
    """Synthetic collated batch with the contract of SortByWidthCollater (reference src/datautils.py:61-176):
    widths sorted descending, image values U[0,1) in the valid region, zero right padding, int32 concatenated
    targets."""
    widths = np.sort(rng.integers(wmin, wmax + 1, size=B))[::-1].astype(np.int32).copy()
    x = np.zeros((B, 1, H, int(widths[0])), np.float32)
    for b in range(B):
        x[b, :, :, :widths[b]] = rng.random((1, H, widths[b]), dtype=np.float32)
    label_lens = np.zeros(B, np.int32)
    labels = []
    for b in range(B):
        t = out_hw(H, int(widths[b]), n_rds)[1]
        L = int(rng.integers(min(lmin, t // 2), min(lmax, t // 2) + 1)) if lmax > 0 else 0
        label_lens[b] = L
        labels.extend(rng.integers(1, n_symbols, size=L).tolist())
    return x, widths, np.array(labels, np.int32), label_lens
