"""Generate the golden fixtures by running the UNMODIFIED reference (imported read-only from /root/reference/src)
on CPU.  Run once in the authoring container: `python tests/golden/make_golden.py`.  The fixtures travel to the
GPU box; /root/reference does not.

Accommodations (no edits to the reference): a stub `textutils` module exporting uxxxx_to_utf8
(the real one needs ICU and absolute data paths), gpu=False/multigpu=False, and fixed `_random_samples` on the two
FractionalMaxPool2d modules so the forward is deterministic.  Weights come from oracle.model_ref.make_state_dict
(numpy PCG64 stream) so they need not be stored.  warp-ctc is not installable: the CTC cost/gradients stored here
come from torch.nn.functional.ctc_loss applied to the REFERENCE's logits (parity with warp-ctc itself is pinned
only by its published known-answer vector, see oracle/ctc_ref.c).
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import model_ref as M  # noqa: E402
from oracle.decode_ref import uxxxx_to_utf8  # noqa: E402

stub = types.ModuleType("textutils")
stub.uxxxx_to_utf8 = uxxxx_to_utf8
sys.modules["textutils"] = stub
sys.path.insert(0, "/root/reference/src")
from alphabet import Alphabet  # noqa: E402
from decoder import ArgmaxDecoder  # noqa: E402
from models.cnnlstm import CnnOcrModel  # noqa: E402

CONFIGS = {
    # name: (hp, n_symbols, B, wmin, wmax, seed)
    "h30": (dict(input_line_height=30, rds_line_height=30, lstm_input_dim=16, num_lstm_layers=2,
                 num_lstm_hidden_units=24, p_lstm_dropout=0.0), 13, 4, 20, 90, 11),
    "h60": (dict(input_line_height=60, rds_line_height=30, lstm_input_dim=24, num_lstm_layers=3,
                 num_lstm_hidden_units=16, p_lstm_dropout=0.0), 29, 3, 60, 150, 12),
    "h120": (dict(input_line_height=120, rds_line_height=30, lstm_input_dim=8, num_lstm_layers=1,
                  num_lstm_hidden_units=8, p_lstm_dropout=0.0), 7, 2, 90, 200, 13),
    # the benchmarked architecture itself (BASELINE cfg2: D128 / 3x512 BiLSTM, h 60 -> rds 30, alphabet 96), small
    # batch: pins H = 512 / K = 1024 / K = 2304 arithmetic to the REAL reference, not to the port
    "cfg2arch": (dict(input_line_height=60, rds_line_height=30, lstm_input_dim=128, num_lstm_layers=3,
                      num_lstm_hidden_units=512, p_lstm_dropout=0.0), 96, 3, 120, 260, 14),
}
# strided slices of the large gradient tensors (a full cnn.21.weight gradient alone would be 2.4 MB)
GRAD_SLICES = {"lstm.weight_hh_l1": (slice(None, None, 64), slice(None, None, 16)),
               "lstm.weight_ih_l2_reverse": (slice(None, None, 64), slice(None, None, 32)),
               "cnn.17.weight": (slice(None, None, 8), slice(None, None, 8)),
               "bridge_layer.0.weight": (slice(None, None, 4), slice(None, None, 16)),
               "prob_layer.0.weight": (slice(None, None, 3), slice(None, None, 16))}
GRAD_KEYS = ["prob_layer.0.bias", "bridge_layer.0.bias", "cnn.0.weight", "cnn.1.weight", "cnn.1.bias",
             "cnn.21.weight", "lstm.bias_hh_l0", "lstm.bias_ih_l0_reverse"]


def alphabet_for(n):
    return Alphabet(["<ctc-blank>"] + ["u%04x" % (0x61 + i) for i in range(n - 1)])


def model_fixture(name):
    hp, A, B, wmin, wmax, seed = CONFIGS[name]
    alpha = alphabet_for(A)
    torch.manual_seed(7)
    model = CnnOcrModel(alphabet=alpha, gpu=False, multigpu=False, verbose=False, **hp)
    sd = M.make_state_dict(hp, A, seed=seed)
    model.load_state_dict(sd, strict=True)
    rng = np.random.default_rng(seed)
    n_rds = M.num_rds_layers(hp["input_line_height"], hp["rds_line_height"])
    x, widths, labels, label_lens = M.synth_batch(rng, B, hp["input_line_height"], wmin, wmax, A, 1, 8, n_rds)
    u1 = rng.random((B, 64, 2)).astype(np.float32)
    u2 = rng.random((B, 128, 2)).astype(np.float32)
    model.cnn[6]._random_samples = torch.from_numpy(u1)
    model.cnn[13]._random_samples = torch.from_numpy(u2)
    out = dict(x=x, widths=widths, labels=labels, label_lens=label_lens, u1=u1, u2=u2, seed=np.int64(seed),
               n_symbols=np.int64(A))
    model.eval()
    with torch.no_grad():
        logits, lens = model(torch.from_numpy(x), torch.from_numpy(widths))
    out["eval_logits"] = logits.numpy()
    out["lens"] = lens.numpy()
    hyp = model.decode_without_lm(logits, lens, uxxxx=True)
    hyp2 = ArgmaxDecoder(alpha).decode(logits, lens, uxxxx=True)
    assert hyp == hyp2
    out["eval_hyp"] = np.array(hyp)
    out["eval_hyp_utf8"] = np.array(model.decode_without_lm(logits, lens, uxxxx=False))
    model.train()
    logits, lens = model(torch.from_numpy(x), torch.from_numpy(widths))
    loss = M.ctc_sum_ref(logits, labels, lens, label_lens)
    loss.backward()
    out["train_logits"] = logits.detach().numpy()
    out["train_loss"] = np.float64(loss.item())
    named = dict(model.named_parameters())
    big = name == "cfg2arch"
    for k in GRAD_KEYS:
        if k in named and not (big and named[k].numel() > 20000):
            out["grad." + k] = named[k].grad.numpy()
    if big:
        for k, sl in GRAD_SLICES.items():
            out["gradslice." + k] = named[k].grad[sl].contiguous().numpy()
            out["gradmax." + k] = np.float64(named[k].grad.abs().max().item())
        out["grad.rapid_ds.00-conv.weight"] = named["rapid_ds.00-conv.weight"].grad.numpy()
        out["grad.lstm.bias_hh_l2"] = named["lstm.bias_hh_l2"].grad.numpy()
    msd = model.state_dict()
    for k in ("cnn.1.running_mean", "cnn.1.running_var", "cnn.21.running_mean", "cnn.21.running_var"):
        out["after." + k] = msd[k].numpy()
    np.savez_compressed(os.path.join(HERE, "model_%s.npz" % name), **out)
    print(name, "logits", logits.shape, "loss", loss.item(), "hyp", hyp[:2])


def decode_fixture():
    rng = np.random.default_rng(5)
    cases = {}
    for A in (5, 97, 120, 121):
        T, B = 41, 5
        x = np.round(rng.normal(size=(T, B, A)) * 2) / 2
        thr = np.float32(3 / A)
        m = rng.random((T, B)) < 0.2
        x[m] = np.minimum(x[m], thr)
        x[rng.random((T, B)) < 0.2, 0] = 7.0
        x = x.astype(np.float32)
        lens = np.array([41, 40, 17, 1, 0], np.int32)
        alpha = alphabet_for(A)
        hyp = ArgmaxDecoder(alpha).decode(torch.from_numpy(x), torch.from_numpy(lens), uxxxx=True)
        hyp8 = ArgmaxDecoder(alpha).decode(torch.from_numpy(x), torch.from_numpy(lens), uxxxx=False)
        cases["A%d.logits" % A] = x
        cases["A%d.lens" % A] = lens
        cases["A%d.hyp" % A] = np.array(hyp)
        cases["A%d.hyp_utf8" % A] = np.array(hyp8)
    np.savez_compressed(os.path.join(HERE, "decode.npz"), **cases)
    print("decode fixture:", [k for k in cases if k.endswith(".hyp")])


if __name__ == "__main__":
    only = sys.argv[1:]
    if not only:
        decode_fixture()
    for name in CONFIGS:
        if not only or name in only:
            model_fixture(name)
