"""GPU parity of the tensor-core conv kernels (csrc/tc_conv.cu) against float64 F.conv2d: forward, data gradient and
weight gradient, incl. ragged tile shapes.  Bound: 1e-5 * max|ref| (fp32 contract)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def close(got, want, what, rtol=1e-5):
    got, want = got.double().cpu(), want.double().cpu()
    err = (got - want).abs().max().item()
    assert err <= rtol * want.abs().max().item(), "%s: err %.3e max %.3e" % (what, err, want.abs().max().item())


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 7, 37, 32, 64), (1, 30, 130, 64, 64), (3, 15, 70, 64, 128),
                                            (2, 7, 294, 256, 256), (1, 5, 33, 128, 32), (2, 4, 9, 32, 256)])
def test_tc_conv_fwd_dgrad_wgrad(cuda, B, H, W, Cin, Cout):
    from vistaocr_b200 import ops
    assert ops.USE_TC
    g = torch.Generator().manual_seed(B + H * 7 + W * 3 + Cin)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / (3.0 * Cin ** 0.5)
    b = torch.randn(Cout, generator=g)
    dz = torch.randn(B, Cout, H, W, generator=g)
    xr, wr = x.double().requires_grad_(True), w.double().requires_grad_(True)
    zr = F.conv2d(xr, wr, b.double(), padding=1)
    zr.backward(dz.double())
    xo, wo, dzo = nhwc(x).to(cuda), w.to(cuda), nhwc(dz).to(cuda)
    z, x_op = ops.conv3x3(xo, wo, b.to(cuda))
    assert x_op is not None  # tensor-core path taken
    close(nchw(z), zr, "fwd")
    dx, dz_op = ops.conv3x3_dgrad(dzo, wo)
    close(nchw(dx), xr.grad, "dgrad")
    dw = ops.conv3x3_wgrad(xo, dzo, x_op, dz_op)
    close(dw, wr.grad, "wgrad")


@pytest.mark.parametrize("xs,ds", [(1e-4, 1e-7), (300.0, 1e3)])
def test_tc_conv_f16_pairs_scaled_operands(cuda, xs, ds):
    """FP16 pair path (Cin, Cout % 64 == 0) with operand magnitudes outside FP16's range."""
    from vistaocr_b200 import ops
    assert ops.USE_F16
    g = torch.Generator().manual_seed(5)
    B, H, W, Cin, Cout = 2, 9, 75, 64, 128
    x = torch.randn(B, Cin, H, W, generator=g) * xs
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / (3.0 * Cin ** 0.5)
    dz = torch.randn(B, Cout, H, W, generator=g) * ds
    xr, wr = x.double().requires_grad_(True), w.double().requires_grad_(True)
    zr = F.conv2d(xr, wr, None, padding=1)
    zr.backward(dz.double())
    xo, wo, dzo = nhwc(x).to(cuda), w.to(cuda), nhwc(dz).to(cuda)
    z, x_op = ops.conv3x3(xo, wo, None)
    close(nchw(z), zr, "fwd")
    dx, dz_op = ops.conv3x3_dgrad(dzo, wo)
    close(nchw(dx), xr.grad, "dgrad")
    close(ops.conv3x3_wgrad(xo, dzo, x_op, dz_op), wr.grad, "wgrad")


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 9, 75, 16, 64), (1, 4, 5, 8, 16), (3, 30, 200, 16, 64)])
def test_narrow_conv_as_patch_gemm(cuda, B, H, W, Cin, Cout):
    """Cin < 64: im2col FP16 pair planes + K = 9*Cin GEMMs (forward and weight gradient)."""
    from vistaocr_b200 import ops
    assert ops._narrow(Cin, Cout)
    g = torch.Generator().manual_seed(11 + W)
    x = torch.randn(B, Cin, H, W, generator=g).relu() * 3.0
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / (3.0 * Cin ** 0.5)
    b = torch.randn(Cout, generator=g)
    dz = torch.randn(B, Cout, H, W, generator=g) * 1e-4
    xr, wr = x.double().requires_grad_(True), w.double().requires_grad_(True)
    zr = F.conv2d(xr, wr, b.double(), padding=1)
    zr.backward(dz.double())
    xo, wo, dzo = nhwc(x).to(cuda), w.to(cuda), nhwc(dz).to(cuda)
    z, x_op = ops.conv3x3(xo, wo, b.to(cuda))
    assert getattr(x_op, "cols", False)
    close(nchw(z), zr, "fwd")
    close(ops.conv3x3_wgrad(xo, dzo, x_op, None), wr.grad, "wgrad")
    close(ops.conv3x3_wgrad(xo, dzo, None, None), wr.grad, "wgrad (patches rebuilt)")
    dx, _ = ops.conv3x3_dgrad(dzo, wo)
    close(nchw(dx), xr.grad, "dgrad")


def test_conv_fwd_same_sign_reduction_keeps_fp32_accuracy(cuda):
    """Worst case for the tensor core's round-toward-zero accumulation: all products positive, K = 9 * 256."""
    from vistaocr_b200 import ops
    g = torch.Generator().manual_seed(2)
    B, H, W, Cin, Cout = 1, 7, 150, 256, 256
    x = torch.rand(B, Cin, H, W, generator=g) + 0.1
    w = torch.rand(Cout, Cin, 3, 3, generator=g) + 0.1
    zr = F.conv2d(x.double(), w.double(), None, padding=1)
    z, _ = ops.conv3x3(nhwc(x).to(cuda), w.to(cuda), None)
    close(nchw(z), zr, "fwd, same-sign K=2304")


def test_wgrad_long_reduction_keeps_fp32_accuracy(cuda):
    """~1.2e5 pixels with a non-zero mean (the worst case for round-toward-zero accumulation)."""
    from vistaocr_b200 import ops
    g = torch.Generator().manual_seed(1)
    B, H, W, Cin, Cout = 8, 30, 500, 64, 64
    x = torch.rand(B, H, W, Cin, generator=g)          # all positive, like post-ReLU activations
    dz = torch.rand(B, H, W, Cout, generator=g) - 0.3
    want = torch.zeros(Cout, Cin, 3, 3, dtype=torch.float64)
    xr = x.double().permute(0, 3, 1, 2).requires_grad_(False)
    wr = torch.zeros(Cout, Cin, 3, 3, dtype=torch.float64, requires_grad=True)
    F.conv2d(xr, wr, None, padding=1).backward(dz.double().permute(0, 3, 1, 2))
    dw = ops.conv3x3_wgrad(x.to(cuda), dz.to(cuda))
    close(dw, wr.grad, "wgrad long K")
    del want


def test_colstats(cuda):
    from vistaocr_b200 import ops
    z = torch.randn(1000, 64, device=cuda) + 0.5
    st = torch.zeros(128, dtype=torch.float64, device=cuda)
    ops.colstats(z, 64, st)
    assert torch.allclose(st[:64], z.double().sum(0), rtol=1e-6)
    assert torch.allclose(st[64:], (z.double() ** 2).sum(0), rtol=1e-6)
