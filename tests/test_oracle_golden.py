"""CPU: the oracle restatements reproduce the golden fixtures generated from the UNMODIFIED reference
(tests/golden/make_golden.py).  Runs anywhere - the fixtures travel, /root/reference does not."""
import os

import numpy as np
import pytest
import torch

from oracle import model_ref as M
from oracle.decode_ref import decode_labels, decode_loop

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CONFIGS = {
    "h30": dict(input_line_height=30, rds_line_height=30, lstm_input_dim=16, num_lstm_layers=2,
                num_lstm_hidden_units=24, p_lstm_dropout=0.0),
    "h60": dict(input_line_height=60, rds_line_height=30, lstm_input_dim=24, num_lstm_layers=3,
                num_lstm_hidden_units=16, p_lstm_dropout=0.0),
    "h120": dict(input_line_height=120, rds_line_height=30, lstm_input_dim=8, num_lstm_layers=1,
                 num_lstm_hidden_units=8, p_lstm_dropout=0.0),
    # the benchmarked architecture (BASELINE cfg2: D128 / 3x512), fixture generated from the REAL reference
    "cfg2arch": dict(input_line_height=60, rds_line_height=30, lstm_input_dim=128, num_lstm_layers=3,
                     num_lstm_hidden_units=512, p_lstm_dropout=0.0),
}
GRAD_SLICES = {"lstm.weight_hh_l1": (slice(None, None, 64), slice(None, None, 16)),
               "lstm.weight_ih_l2_reverse": (slice(None, None, 64), slice(None, None, 32)),
               "cnn.17.weight": (slice(None, None, 8), slice(None, None, 8)),
               "bridge_layer.0.weight": (slice(None, None, 4), slice(None, None, 16)),
               "prob_layer.0.weight": (slice(None, None, 3), slice(None, None, 16))}


@pytest.mark.parametrize("A", [5, 97, 120, 121])
def test_decode_oracle_matches_reference_fixture(A):
    z = np.load(os.path.join(GOLD, "decode.npz"))
    x, lens = z["A%d.logits" % A], z["A%d.lens" % A]
    idx_to_char = {i: ("<ctc-blank>" if i == 0 else "u%04x" % (0x61 + i - 1)) for i in range(A)}
    assert decode_loop(x, lens, idx_to_char, uxxxx=True) == z["A%d.hyp" % A].tolist()
    assert decode_loop(x, lens, idx_to_char, uxxxx=False) == z["A%d.hyp_utf8" % A].tolist()
    labs, _ = decode_labels(x, lens, A)
    assert [" ".join(idx_to_char[k] for k in l) for l in labs] == z["A%d.hyp" % A].tolist()


@pytest.mark.parametrize("name", ["h30", "h60", "h120", "cfg2arch"])
def test_model_oracle_matches_reference_fixture(name):
    hp = CONFIGS[name]
    z = np.load(os.path.join(GOLD, "model_%s.npz" % name))
    A = int(z["n_symbols"])
    sd = M.make_state_dict(hp, A, seed=int(z["seed"]))
    x = torch.from_numpy(z["x"])
    u = (torch.from_numpy(z["u1"]), torch.from_numpy(z["u2"]))
    with torch.no_grad():
        logits, lens = M.forward_ref(sd, x, z["widths"], hp, u, training=False)
    assert lens.tolist() == z["lens"].tolist()
    assert np.abs(logits.numpy() - z["eval_logits"]).max() <= 2e-6  # same torch CPU kernels: equal up to threading
    idx_to_char = {i: ("<ctc-blank>" if i == 0 else "u%04x" % (0x61 + i - 1)) for i in range(A)}
    assert decode_loop(logits.numpy(), lens.numpy(), idx_to_char, uxxxx=True) == z["eval_hyp"].tolist()
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running" not in k}
    state = dict(sd)
    state.update(params)
    upd = {}
    logits, lens = M.forward_ref(state, x, z["widths"], hp, u, training=True, bn_updates=upd)
    assert np.abs(logits.detach().numpy() - z["train_logits"]).max() <= 2e-6
    loss = M.ctc_sum_ref(logits, z["labels"], lens, z["label_lens"])
    assert abs(loss.item() - float(z["train_loss"])) <= 1e-5 * float(z["train_loss"])
    loss.backward()
    for k in z.files:
        if k.startswith("grad."):
            assert np.abs(params[k[5:]].grad.numpy() - z[k]).max() <= 1e-4 * np.abs(z[k]).max() + 1e-6, k
        if k.startswith("gradslice."):
            got = params[k[10:]].grad[GRAD_SLICES[k[10:]]].numpy()
            assert np.abs(got - z[k]).max() <= 1e-4 * float(z["gradmax." + k[10:]]) + 1e-6, k
        if k.startswith("after."):
            assert np.abs(upd[k[6:]].numpy() - z[k]).max() <= 1e-6, k


def test_output_size_calculator_traps():
    # reference cnn_input_size_to_output_size: floor(floor(w*0.7)*0.7) in float64
    assert M.out_hw(30, 350, 0) == (7, 170)  # floor(350*0.7)=244 (not 245), floor(244*0.7)=170
    assert [M.out_hw(30, w, 0)[1] for w in (15, 200, 400, 800)] == [7, 98, 196, 392]
    assert M.out_hw(60, 800, 1)[1] == 196 and M.out_hw(120, 2000, 2)[1] == 244
    from vistaocr_b200.ops import out_hw
    for h, n in ((30, 0), (60, 1), (120, 2)):
        for w in range(16, 2100, 7):
            assert out_hw(h, w, n) == M.out_hw(h, w, n)
