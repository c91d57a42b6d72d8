"""GPU: the reduced-precision mode of the tensor-core kernels (vistaocr_b200.set_precision("fp16"): one product on the
FP16 hi planes, fp32 accumulation - BASELINE.json cfg3's "bf16 training" point, with 11 instead of 8 mantissa bits).
Documented bounds, relative to the tensor's max: single GEMM / convolution 1e-3; logits of the whole network 1e-2; CTC loss
1e-3 relative; parameter gradients stay aligned with the float64 gradients (cosine > 0.99; measured 0.9977 for the
first convolution, the parameter farthest from the loss, > 0.9995 elsewhere).  The default mode must be
untouched afterwards (the switch is process-wide)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import model_ref as M

pytestmark = pytest.mark.gpu


@pytest.fixture
def fp16_mode(cuda):
    import vistaocr_b200
    prev = vistaocr_b200.set_precision("fp16")
    assert vistaocr_b200.get_precision() == "fp16"
    yield cuda
    vistaocr_b200.set_precision(prev)
    assert vistaocr_b200.get_precision() == "fp32"


def _rel(got, want):
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    return ((got - want).abs().max() / want.abs().max()).item()


@pytest.mark.parametrize("M_,N_,K_", [(300, 256, 1024), (18816, 128, 1792), (512, 2048, 128), (2048, 1024, 18816)])
def test_gemm_fp16_operands(fp16_mode, M_, N_, K_):
    from vistaocr_b200 import ops
    g = torch.Generator().manual_seed(M_ + N_)
    A = torch.randn(M_, K_, generator=g)
    B = torch.randn(N_, K_, generator=g) / K_ ** 0.5
    C = torch.empty(M_, N_, device=fp16_mode)
    ops.mm(0, 1, M_, N_, K_, ops.Operand(A.to(fp16_mode)), K_, ops.Operand(B.to(fp16_mode)), K_, C, N_)
    err = _rel(C, A.double() @ B.double().t())
    assert 1e-6 < err <= 1e-3, err  # really the single-product path, and inside its bound


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 7, 294, 256, 256), (3, 15, 70, 64, 128), (2, 30, 130, 64, 64)])
def test_conv_fp16_operands(fp16_mode, B, H, W, Cin, Cout):
    from vistaocr_b200 import ops
    g = torch.Generator().manual_seed(B + H + Cin)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / (3.0 * Cin ** 0.5)
    dz = torch.randn(B, Cout, H, W, generator=g)
    xr, wr = x.double().requires_grad_(True), w.double().requires_grad_(True)
    zr = F.conv2d(xr, wr, None, padding=1)
    zr.backward(dz.double())
    nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous()
    nchw = lambda t: t.permute(0, 3, 1, 2).contiguous()
    xo, wo, dzo = nhwc(x).to(fp16_mode), w.to(fp16_mode), nhwc(dz).to(fp16_mode)
    z, x_op = ops.conv3x3(xo, wo, None)
    assert 1e-6 < _rel(nchw(z), zr) <= 1e-3
    dx, _ = ops.conv3x3_dgrad(dzo, wo)
    assert _rel(nchw(dx), xr.grad) <= 1e-3
    dw = ops.conv3x3_wgrad(xo, dzo, x_op=x_op)
    assert _rel(dw, wr.grad) <= 1e-3


def test_whole_path_fp16_operands(fp16_mode):
    from vistaocr_b200 import Alphabet, CnnOcrModel, CTCLoss
    hp = dict(input_line_height=30, rds_line_height=30, lstm_input_dim=32, num_lstm_layers=2,
              num_lstm_hidden_units=40, p_lstm_dropout=0.0)
    A = 31
    sd = M.make_state_dict(hp, A, seed=5)
    model = CnnOcrModel(alphabet=Alphabet(["<ctc-blank>"] + ["u%04x" % (0x61 + i) for i in range(A - 1)]), verbose=False, **hp)
    model.load_state_dict(sd, strict=True)
    rng = np.random.default_rng(9)
    x, widths, labels, label_lens = M.synth_batch(rng, 5, 30, 40, 160, A, 2, 10)
    u1 = torch.from_numpy(rng.random((5, 64, 2)).astype(np.float32))
    u2 = torch.from_numpy(rng.random((5, 128, 2)).astype(np.float32))
    model.cnn[6]._random_samples, model.cnn[13]._random_samples = u1, u2
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    for k in sd64:
        if sd64[k].is_floating_point() and "running" not in k:
            sd64[k].requires_grad_(True)
    model.train()
    logits, lens = model(torch.from_numpy(x).to(fp16_mode), torch.from_numpy(widths))
    want, wlens = M.forward_ref(sd64, torch.from_numpy(x).double(), widths, hp, (u1, u2), training=True,
                                bn_updates={}, use_nn_lstm=False)
    err = _rel(logits, want)
    assert 1e-6 < err <= 1e-2, err
    loss = CTCLoss()(logits, torch.from_numpy(labels), lens, torch.from_numpy(label_lens))
    wloss = M.ctc_sum_ref(want, labels, wlens, label_lens)
    assert abs(loss.data[0].item() - wloss.item()) <= 1e-3 * abs(wloss.item())
    loss.backward()
    wloss.backward()
    for k, p in model.named_parameters():
        w = sd64[k].grad
        if k.startswith("cnn.") and k.endswith(".bias") and int(k.split(".")[1]) in M.CONV_IDX:
            continue  # mathematically zero (conv bias before train-mode BN)
        a, b = p.grad.double().cpu().flatten(), w.flatten()
        cos = (a @ b / (a.norm() * b.norm() + 1e-300)).item()
        assert cos > 0.99, (k, cos)
