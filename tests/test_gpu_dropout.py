"""GPU parity of the inter-layer LSTM dropout (reference nn.LSTM(dropout=p), cnnlstm.py:148-149, p = 0.5 at
train_cnn_lstm.py:331): the in-kernel Philox mask bit for bit against oracle/philox_ref.py, the op and its backward,
and the WHOLE model in training mode at p = 0.5 against the oracle run with the same masks - both with injected masks
and with the in-kernel stream."""
import numpy as np
import pytest
import torch

from oracle import model_ref as M
from oracle.philox_ref import dropout_ref, keep_mask

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,p,seed,offset", [(1, 0.5, 0, 0), (7, 0.5, 1, 2), (4096, 0.5, 1234, 77),
                                              (100_003, 0.3, 2 ** 62 + 5, 2 ** 33 + 9), (1 << 20, 0.9, 7, 1)])
def test_mask_matches_philox_oracle(cuda, n, p, seed, offset):
    from vistaocr_b200 import ops
    got = ops.dropout_keep_mask(n, p, seed, offset, cuda).cpu().numpy()
    assert np.array_equal(got, keep_mask(n, p, seed, offset))


def test_dropout_forward_backward_and_rng_state(cuda):
    from vistaocr_b200 import ops
    rng = np.random.default_rng(0)
    x = rng.normal(size=(37, 5, 11)).astype(np.float32)  # 2035 elements: exercises the scalar tail
    xt = torch.from_numpy(x).to(cuda).requires_grad_(True)
    state = ops.new_rng_state(cuda, seed=99, offset=1000)
    y = ops.dropout(xt, 0.5, rng=state, site=3)
    mask = keep_mask(x.size, 0.5, 99, 1003)
    assert np.array_equal(y.detach().cpu().numpy(), dropout_ref(x, 0.5, mask))
    ops.rng_advance(state, 2)  # the state moves on before backward runs: backward must still use the forward's mask
    assert state.tolist() == [99, 1002]
    dy = rng.normal(size=x.shape).astype(np.float32)
    y.backward(torch.from_numpy(dy).to(cuda))
    assert np.array_equal(xt.grad.cpu().numpy(), dropout_ref(dy, 0.5, mask))
    y2 = ops.dropout(xt.detach(), 0.5, rng=state, site=3)  # next step: offset 1005, a different mask
    assert np.array_equal(y2.cpu().numpy(), dropout_ref(x, 0.5, keep_mask(x.size, 0.5, 99, 1005)))
    # injected mask
    inj = (rng.random(x.shape) < 0.7).astype(np.uint8)
    x3 = torch.from_numpy(x).to(cuda).requires_grad_(True)
    y3 = ops.dropout(x3, 0.25, mask=torch.from_numpy(inj))
    assert np.array_equal(y3.detach().cpu().numpy(), dropout_ref(x, 0.25, inj))
    y3.backward(torch.from_numpy(dy).to(cuda))
    assert np.array_equal(x3.grad.cpu().numpy(), dropout_ref(dy, 0.25, inj))
    with pytest.raises(Exception):
        ops.dropout(x3, 0.5)  # neither a stream nor a mask


def _alphabet(n):
    from vistaocr_b200 import Alphabet
    return Alphabet(["<ctc-blank>"] + ["u%04x" % (0x61 + i) for i in range(n - 1)])


@pytest.mark.parametrize("mode", ["injected", "philox"])
def test_model_training_with_dropout_matches_oracle(cuda, mode):
    """p = 0.5 between three LSTM layers: logits, CTC loss and every parameter gradient against the float64 oracle that
    multiplies the same masks in (oracle/model_ref.py::bilstm_ref(dropout_masks=...))."""
    from vistaocr_b200 import CnnOcrModel, CTCLoss, ops
    hp = dict(input_line_height=30, rds_line_height=30, lstm_input_dim=32, num_lstm_layers=3,
              num_lstm_hidden_units=40, p_lstm_dropout=0.5)
    A, B = 23, 5
    sd = M.make_state_dict(hp, A, seed=15)
    model = CnnOcrModel(alphabet=_alphabet(A), verbose=False, **hp)
    model.load_state_dict(sd, strict=True)
    rng = np.random.default_rng(4)
    x, widths, labels, label_lens = M.synth_batch(rng, B, 30, 40, 170, A, 2, 10)
    u1 = torch.from_numpy(rng.random((B, 64, 2)).astype(np.float32))
    u2 = torch.from_numpy(rng.random((B, 128, 2)).astype(np.float32))
    model.cnn[6]._random_samples, model.cnn[13]._random_samples = u1, u2
    lens = [M.out_hw(30, int(w), 0)[1] for w in widths]
    tmax, wf, H2 = max(lens), M.out_hw(30, int(widths[0]), 0)[1], 2 * hp["num_lstm_hidden_units"]
    n = tmax * B * H2
    if mode == "injected":
        keep = [(rng.random((tmax, B, H2)) < 0.5).astype(np.uint8) for _ in range(2)]
        model._dropout_masks = [torch.from_numpy(k) for k in keep]
    else:
        model.set_dropout_seed(4242, offset=10)
        keep = [keep_mask(n, 0.5, 4242, 10 + l).reshape(tmax, B, H2) for l in range(2)]
    # eval mode ignores dropout entirely (checked first: the training forward below updates the running statistics)
    model.eval()
    with torch.no_grad():
        ev, _ = model(torch.from_numpy(x).to(cuda), torch.from_numpy(widths))
    sd64e = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    wev, _ = M.forward_ref(sd64e, torch.from_numpy(x).double(), widths, hp, (u1, u2), training=False, use_nn_lstm=False)
    assert (ev.double().cpu() - wev.detach()).abs().max().item() <= 2e-5 * wev.abs().max().item() + 2e-6
    model.train()
    logits, olens = model(torch.from_numpy(x).to(cuda), torch.from_numpy(widths))
    if mode == "philox":
        assert model._dropout_rng.tolist() == [4242, 12]  # advanced by L-1 sites
        for l in range(2):  # and the kernel's own view of the stream is the oracle's
            assert np.array_equal(ops.dropout_keep_mask(n, 0.5, 4242, 10 + l, cuda).cpu().numpy(), keep[l].reshape(-1))
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    for k in sd64:
        if sd64[k].is_floating_point() and "running" not in k:
            sd64[k].requires_grad_(True)
    masks64 = []
    for k in keep:  # oracle masks: scaled by 1/(1-p), padded to the CNN's frame count
        m = torch.ones((wf, B, H2), dtype=torch.float64)
        m[:tmax] = torch.from_numpy(k).double() * 2.0
        masks64.append(m)
    want, wlens = M.forward_ref(sd64, torch.from_numpy(x).double(), widths, hp, (u1, u2), training=True,
                                bn_updates={}, dropout_masks=masks64, use_nn_lstm=False)
    assert olens.tolist() == wlens.tolist() and logits.shape == want.shape
    err = (logits.detach().double().cpu() - want.detach()).abs().max().item()
    assert err <= 2e-5 * want.abs().max().item() + 2e-6, err
    # the masks matter: without them the logits are far away (guards against a silently skipped dropout)
    plain, _ = M.forward_ref(sd64, torch.from_numpy(x).double(), widths, hp, (u1, u2), training=True, bn_updates={},
                             use_nn_lstm=False)
    assert (plain.detach() - want.detach()).abs().max().item() > 1e-2
    loss = CTCLoss()(logits, torch.from_numpy(labels), olens, torch.from_numpy(label_lens))
    wloss = M.ctc_sum_ref(want, labels, wlens, label_lens)
    assert abs(loss.data[0].item() - wloss.item()) <= 2e-5 * abs(wloss.item())
    loss.backward()
    wloss.backward()
    # Bounds from profiles/r02_parity_errors.md (12 seeds of this very configuration, CUDA path and the reference's own
    # fp32 arithmetic, both against float64).  Everything downstream of the CNN: ours 2e-6..5e-6 (fp32 oracle: 3e-5..4e-5)
    # -> 5e-5.  CNN parameters: a BatchNorm -> ReLU / max-pool decision that sits within rounding of a tie flips between
    # ANY two fp32 evaluations and moves a conv gradient by whole terms - both implementations show up to 3e-2 of the
    # tensor's max there (2e-2 normwise), so that is the honest bound for a fresh random batch; the kernels themselves
    # are held to 1e-5-level bounds by the per-op tests (test_gpu_ops / test_gpu_tc_conv) and the fixed-seed fixtures.
    num = na = nb = 0.0
    for k, p in model.named_parameters():
        if k.startswith("cnn.") and k.endswith(".bias") and int(k.split(".")[1]) in M.CONV_IDX:
            continue
        w = sd64[k].grad
        d = p.grad.double().cpu() - w
        e, l2 = d.abs().max().item() / w.abs().max().item(), (d.norm() / w.norm()).item()
        if k.startswith("cnn.") or k.startswith("rapid_ds."):
            assert e <= 6e-2 and l2 <= 4e-2, (k, e, l2)
        else:
            assert e <= 5e-5 and l2 <= 5e-5, (k, e, l2)
        num += (p.grad.double().cpu() * w).sum().item()
        na += (p.grad.double().cpu() ** 2).sum().item()
        nb += (w ** 2).sum().item()
    assert num / (na ** 0.5 * nb ** 0.5) >= 1.0 - 1e-4  # the whole gradient, normwise
