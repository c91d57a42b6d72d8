"""GPU parity of the individual compute kernels (through the C ABI) against plain PyTorch CPU float64 references of
the same op.  fp32 bound per tensor: max|ours - ref| <= 1e-5 * max|ref| (RTOL below); gradients
that are mathematically ~0 are checked against an absolute bound."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import model_ref as M

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def close(got, want, rtol=RTOL, atol=0.0, what=""):
    got = got.detach().double().cpu()
    want = want.detach().double().cpu()
    assert got.shape == want.shape, (what, got.shape, want.shape)
    err = (got - want).abs().max().item()
    bound = rtol * want.abs().max().item() + atol
    assert err <= bound, "%s: max err %.3e > bound %.3e" % (what, err, bound)


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


@pytest.mark.parametrize("ta,tb,M_,N_,K_,bias,relu", [(0, 1, 300, 128, 1792, True, True), (0, 1, 257, 97, 70, True, False),
                                                      (0, 0, 130, 70, 97, False, False), (1, 0, 97, 130, 257, False, False),
                                                      (1, 1, 33, 65, 17, True, False), (0, 1, 5, 3, 1, False, False),
                                                      (0, 1, 1000, 2048, 128, True, False)])
def test_gemm(cuda, ta, tb, M_, N_, K_, bias, relu):
    from vistaocr_b200 import ops
    g = torch.Generator().manual_seed(M_ + N_ + K_)
    A = torch.randn((K_, M_) if ta else (M_, K_), generator=g)
    B = torch.randn((N_, K_) if tb else (K_, N_), generator=g)
    bvec = torch.randn(N_, generator=g) if bias else None
    C = torch.full((M_, N_), 7.0, device=cuda)
    ops.gemm(ta, tb, M_, N_, K_, A.to(cuda), A.shape[1], B.to(cuda), B.shape[1], C, N_,
             bias=bvec.to(cuda) if bias else None, relu=relu)
    want = (A.double().t() if ta else A.double()) @ (B.double().t() if tb else B.double())
    if bias:
        want = want + bvec.double()
    if relu:
        want = want.relu()
    close(C, want, what="gemm")
    # accumulate + strided C
    C2 = torch.ones((M_, N_ + 3), device=cuda)
    ops.gemm(ta, tb, M_, N_, K_, A.to(cuda), A.shape[1], B.to(cuda), B.shape[1], C2, N_ + 3, accumulate=True)
    want2 = (A.double().t() if ta else A.double()) @ (B.double().t() if tb else B.double()) + 1
    close(C2[:, :N_], want2, what="gemm accumulate")
    assert (C2[:, N_:] == 1).all()


@pytest.mark.parametrize("M,K,N", [(70, 50, 33), (300, 64, 166), (130, 1024, 97), (257, 128, 6)])
def test_linear_autograd(cuda, M, K, N):
    """(N not a multiple of 8 with K % 8 == 0: the backward GEMMs run on the tensor cores over zero-padded dy / W -
    the alphabet of 166 symbols of cfg3.)"""
    from vistaocr_b200 import ops
    g = torch.Generator().manual_seed(M + N)
    x = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g)
    b = torch.randn(N, generator=g)
    dy = torch.randn(M, N, generator=g)
    xs = [t.clone().to(cuda).requires_grad_(True) for t in (x, w, b)]
    y = ops.linear(*xs, relu=True)
    y.backward(dy.to(cuda))
    xr = [t.clone().double().requires_grad_(True) for t in (x, w, b)]
    yr = F.relu(F.linear(*xr))
    yr.backward(dy.double())
    close(y, yr, what="linear")
    for a, r, n in zip(xs, xr, "xwb"):
        close(a.grad, r.grad, what="linear d" + n)


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 9, 37, 1, 64), (3, 7, 21, 64, 64), (2, 15, 50, 64, 128),
                                            (1, 7, 33, 128, 256), (2, 5, 140, 16, 64), (1, 3, 5, 256, 256)])
@pytest.mark.parametrize("training", [True, False])
def test_conv_bn_relu_block(cuda, B, H, W, Cin, Cout, training):
    from vistaocr_b200 import ops
    g = torch.Generator().manual_seed(B * 1000 + W + Cin)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / (3.0 * Cin ** 0.5)
    b = torch.randn(Cout, generator=g) * 0.1
    gamma = 1 + 0.3 * torch.randn(Cout, generator=g)
    beta = 0.3 * torch.randn(Cout, generator=g)
    rm = 0.1 * torch.randn(Cout, generator=g)
    rv = 1 + 0.3 * torch.rand(Cout, generator=g)
    dy = torch.randn(B, Cout, H, W, generator=g)
    # reference, float64
    xr, wr, br, gr, ber = [t.clone().double().requires_grad_(True) for t in (x, w, b, gamma, beta)]
    rmr, rvr = rm.clone().double(), rv.clone().double()
    z = F.conv2d(xr, wr, br, padding=1)
    yr = F.relu(F.batch_norm(z, rmr, rvr, gr, ber, training, 0.1, 1e-5))
    yr.backward(dy.double())
    # ours
    xo = nhwc(x).to(cuda).requires_grad_(True)
    wo, bo, go, beo = [t.clone().to(cuda).requires_grad_(True) for t in (w, b, gamma, beta)]
    rmo, rvo = rm.clone().to(cuda), rv.clone().to(cuda)
    yo = ops.conv_bn_relu(xo, wo, bo, go, beo, rmo, rvo, training)
    yo.backward(nhwc(dy).to(cuda))
    close(nchw(yo), yr, what="conv_bn_relu fwd")
    close(rmo, rmr, what="running_mean")
    close(rvo, rvr, what="running_var")
    close(nchw(xo.grad), xr.grad, what="dx")
    close(wo.grad, wr.grad, what="dw")
    close(go.grad, gr.grad, what="dgamma")
    close(beo.grad, ber.grad, what="dbeta")
    if training:  # the conv bias in front of a train-mode BN has a mathematically zero gradient (rounding noise)
        assert bo.grad.abs().max().item() <= 1e-3 * max(1.0, wr.grad.abs().max().item())
    else:
        close(bo.grad, br.grad, what="dbias")
    # sequence layout of the last block: [W, B, H*C] with feature y*C + c
    xo2 = nhwc(x).to(cuda)
    ys = ops.conv_bn_relu(xo2, w.to(cuda), b.to(cuda), gamma.to(cuda), beta.to(cuda), rm.clone().to(cuda),
                          rv.clone().to(cuda), training, seq_layout=True)
    want = yr.detach().permute(3, 0, 2, 1).reshape(W, B, H * Cout)
    close(ys, want, what="seq layout")


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(3, 7, 21, 64, 64), (2, 15, 50, 64, 128), (1, 7, 133, 128, 256),
                                            (2, 3, 70, 256, 256)])
def test_conv_bn_relu_eval_fused_epilogue(cuda, B, H, W, Cin, Cout):
    """Inference: conv + BN(running stats) + ReLU in one kernel (vocr_tc_conv3x3_bnrelu_f16) - the fp32 activation is bit
    for bit what conv -> z -> bn_relu_apply writes, the FP16 pair planes it emits for the next conv decode to the same
    values to 22 bits of their (analytic, loose) bound, a two-block chain through the planes-only hand-over meets the
    float64 reference, and so does the time-major sequence layout of the last block."""
    from vistaocr_b200 import ops
    g = torch.Generator().manual_seed(B * 77 + W + Cout)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / (3.0 * Cin ** 0.5)
    b = torch.randn(Cout, generator=g) * 0.1
    gamma, beta = 1 + 0.3 * torch.randn(Cout, generator=g), 0.3 * torch.randn(Cout, generator=g)
    rm, rv = 0.1 * torch.randn(Cout, generator=g), 1 + 0.3 * torch.rand(Cout, generator=g)
    w2 = torch.randn(64, Cout, 3, 3, generator=g) / (3.0 * Cout ** 0.5)
    g2, b2 = 1 + 0.3 * torch.randn(64, generator=g), 0.3 * torch.randn(64, generator=g)
    rm2, rv2 = 0.1 * torch.randn(64, generator=g), 1 + 0.3 * torch.rand(64, generator=g)
    y1 = F.relu(F.batch_norm(F.conv2d(x.double(), w.double(), b.double(), padding=1), rm.double(), rv.double(),
                             gamma.double(), beta.double(), False, 0.1, 1e-5))
    y2 = F.relu(F.batch_norm(F.conv2d(y1, w2.double(), None, padding=1), rm2.double(), rv2.double(), g2.double(),
                             b2.double(), False, 0.1, 1e-5))
    dev = [t.to(cuda) for t in (w, b, gamma, beta, rm, rv)]
    dev2 = [w2.to(cuda), None, g2.to(cuda), b2.to(cuda), rm2.to(cuda), rv2.to(cuda)]
    xo = nhwc(x).to(cuda)
    assert ops.FUSE_EVAL
    with torch.no_grad():
        fused = ops.conv_bn_relu(xo, *dev, False)
        ops.FUSE_EVAL = False
        try:
            plain = ops.conv_bn_relu(xo, *dev, False)
        finally:
            ops.FUSE_EVAL = True
        assert torch.equal(fused, plain)
        close(nchw(fused), y1, what="fused eval block")
        if Cout % 64 == 0:
            hi, lo, st = fused._vocr_op.split16()
            e = int(st[0].item())
            dec = (hi.double() + lo.double() / 2048.0) * 2.0 ** (-e)
            bound = float(fused._vocr_plane_bound[0].item())
            assert bound >= fused.abs().max().item() == float(fused._vocr_bound[0].item())
            assert (dec - fused.double()).abs().max().item() <= bound * 2.0 ** -21
            # chain: the first block hands over planes only, the second reads them
            a1 = ops.conv_bn_relu(xo, *dev, False, allow_planes_only=True)
            a2 = ops.conv_bn_relu(a1, *dev2, False)
            close(nchw(a2), y2, what="two fused blocks")
        ys = ops.conv_bn_relu(xo, *dev, False, seq_layout=True)
        close(ys, y1.permute(3, 0, 2, 1).reshape(W, B, H * Cout), what="fused eval block, seq layout")


@pytest.mark.parametrize("B,H,W,Cin", [(2, 12, 38, 1), (1, 8, 22, 16), (2, 6, 9, 3)])
def test_rapid_ds(cuda, B, H, W, Cin):
    from vistaocr_b200 import ops
    g = torch.Generator().manual_seed(H + W)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(16, Cin, 3, 3, generator=g) / 3.0
    b = torch.randn(16, generator=g) * 0.2
    dy = torch.randn(B, 16, H // 2, W // 2, generator=g)
    xr, wr, br = [t.clone().double().requires_grad_(True) for t in (x, w, b)]
    yr = F.max_pool2d(F.relu(F.conv2d(xr, wr, br, padding=1)), 2, stride=2)
    yr.backward(dy.double())
    xo = nhwc(x).to(cuda).requires_grad_(True)
    wo, bo = [t.clone().to(cuda).requires_grad_(True) for t in (w, b)]
    yo = ops.rapid_ds(xo, wo, bo)
    yo.backward(nhwc(dy).to(cuda))
    close(nchw(yo), yr, what="rds fwd")
    close(nchw(xo.grad), xr.grad, what="rds dx")
    close(wo.grad, wr.grad, what="rds dw")
    close(bo.grad, br.grad, what="rds db")


@pytest.mark.parametrize("B,H,W,C", [(3, 30, 77, 64), (2, 15, 53, 128), (1, 30, 800, 64), (2, 15, 560, 128),
                                     (2, 4, 5, 8)])
def test_fracpool_bit_exact(cuda, B, H, W, C):
    """Window selection must equal ATen's for the same samples: outputs are bit-exact, not just close."""
    from vistaocr_b200 import ops
    g = torch.Generator().manual_seed(W)
    x = torch.randn(B, C, H, W, generator=g)
    u = torch.rand(B, C, 2, generator=g)
    dy = torch.randn(B, C, int(H * 0.5), int(W * 0.7), generator=g)
    xr = x.clone().requires_grad_(True)
    yr = F.fractional_max_pool2d(xr, 2, output_ratio=(0.5, 0.7), _random_samples=u)
    yr.backward(dy)
    xo = nhwc(x).to(cuda).requires_grad_(True)
    yo = ops.fracpool(xo, u.to(cuda))
    yo.backward(nhwc(dy).to(cuda))
    assert torch.equal(nchw(yo).cpu(), yr.detach())
    close(nchw(xo.grad), xr.grad, rtol=1e-6, what="fracpool dx")
    # the backward scatter is deterministic (four race-free parity passes, no atomics): bit-identical reruns
    g1 = xo.grad.clone()
    xo.grad = None
    ops.fracpool(xo, u.to(cuda)).backward(nhwc(dy).to(cuda))
    assert torch.equal(g1, xo.grad)
    # and the oracle's own restatement of the interval rule
    assert torch.equal(M.fmp_ref(x[:1, :4], u[:1, :4]), yr.detach()[:1, :4])


@pytest.mark.parametrize("T,B,D,H,ragged", [(9, 3, 5, 8, True), (20, 5, 16, 24, True), (7, 33, 12, 16, True),
                                            (12, 2, 8, 40, False), (6, 70, 4, 8, True), (5, 4, 128, 512, True),
                                            (12, 20, 16, 256, True), (8, 18, 8, 500, True), (6, 5, 8, 200, True),
                                            # batches >= 128 use 64-sample cluster work items (W_lo split TMEM / smem)
                                            (7, 130, 12, 512, True), (6, 200, 8, 72, True), (5, 448, 4, 264, True)])
def test_bilstm_layer(cuda, T, B, D, H, ragged):
    from vistaocr_b200 import ops
    g = torch.Generator().manual_seed(T * 100 + B)
    x = torch.randn(T, B, D, generator=g)
    lens = sorted([int(v) for v in torch.randint(1, T + 1, (B,), generator=g)], reverse=True) if ragged else [T] * B
    lens[0] = T
    k = 1.0 / H ** 0.5
    P = {n: (torch.rand(s, generator=g) * 2 - 1) * k for n, s in
         (("w_ih_f", (4 * H, D)), ("w_hh_f", (4 * H, H)), ("b_ih_f", (4 * H,)), ("b_hh_f", (4 * H,)),
          ("w_ih_r", (4 * H, D)), ("w_hh_r", (4 * H, H)), ("b_ih_r", (4 * H,)), ("b_hh_r", (4 * H,)))}
    dy = torch.randn(T, B, 2 * H, generator=g)
    # reference: explicit masked recurrence, float64, autograd
    R = {n: v.clone().double().requires_grad_(True) for n, v in P.items()}
    xr = x.clone().double().requires_grad_(True)
    of = M.lstm_cell_loop(xr, lens, R["w_ih_f"], R["w_hh_f"], R["b_ih_f"], R["b_hh_f"], False)
    orv = M.lstm_cell_loop(xr, lens, R["w_ih_r"], R["w_hh_r"], R["b_ih_r"], R["b_hh_r"], True)
    yr = torch.cat([of, orv], 2)
    yr.backward(dy.double())
    # ours
    G = {n: v.clone().to(cuda).requires_grad_(True) for n, v in P.items()}
    xo = x.clone().to(cuda).requires_grad_(True)
    w_ih = torch.cat([G["w_ih_f"], G["w_ih_r"]], 0)
    w_hh = torch.stack([G["w_hh_f"], G["w_hh_r"]], 0)
    bias = torch.cat([G["b_ih_f"] + G["b_hh_f"], G["b_ih_r"] + G["b_hh_r"]])
    lens_dev = torch.tensor(lens, dtype=torch.int32, device=cuda)
    yo = ops.bilstm_layer(xo, w_ih, w_hh, bias, lens_dev, max(lens))
    yo.backward(dy.to(cuda))
    close(yo, yr, what="lstm out")
    for b in range(B):
        assert not yo[lens[b]:, b].any()  # exact zeros beyond the sample's length
    close(xo.grad, xr.grad, what="lstm dx")
    for n in P:
        close(G[n].grad, R[n].grad, what="lstm d" + n)


@pytest.mark.parametrize("T,B,D,H", [(9, 40, 8, 512), (11, 64, 8, 40), (6, 5, 8, 200), (7, 33, 12, 16)])
def test_bilstm_layer_cluster_backward(cuda, monkeypatch, T, B, D, H):
    """The opt-in cluster-resident tcgen05 backward recurrence (csrc/lstm_tc.cu, VOCR_LSTM_TC_BWD=1; measured slower than
    the default kernel, kept as a checked alternative) meets the same bounds."""
    monkeypatch.setenv("VOCR_LSTM_TC_BWD", "1")
    test_bilstm_layer(cuda, T, B, D, H, True)


def test_clamp_adam_matches_torch(cuda):
    from vistaocr_b200 import ops
    g = torch.Generator().manual_seed(3)
    n = 1003
    p0 = torch.randn(n, generator=g)
    pr = p0.clone().double().requires_grad_(True)
    opt = torch.optim.Adam([pr], lr=1e-3, weight_decay=0.01)
    po, m, v = p0.clone().to(cuda), torch.zeros(n, device=cuda), torch.zeros(n, device=cuda)
    for step in range(1, 6):
        grad = torch.randn(n, generator=g) * 4  # some entries beyond the +-5 clamp
        pr.grad = grad.double().clamp(-5, 5)
        opt.step()
        ops.clamp_adam_step(po, grad.to(cuda), m, v, step, lr=1e-3, weight_decay=0.01, clamp=5.0)
        close(po, pr, rtol=2e-6, what="adam step %d" % step)


def test_clamp_adam_state_dict_round_trip_and_adam_compat(cuda):
    """The reference snapshots `optimizer.state_dict()` (train_cnn_lstm.py:427-438): a resumed run must continue with
    the same moments and step count, and the format is torch.optim.Adam's (both directions)."""
    from vistaocr_b200 import ClampAdam
    g = torch.Generator().manual_seed(5)
    shapes = [(7, 3), (5,), (2, 3, 3)]
    init = [torch.randn(s, generator=g) for s in shapes]
    grads = [[torch.randn(s, generator=g) * 3 for s in shapes] for _ in range(5)]

    def params():
        return [torch.nn.Parameter(t.clone().to(cuda)) for t in init]

    def run(opt, ps, steps):
        for gs in steps:
            opt.zero_grad()
            for p, gr in zip(ps, gs):
                if p.grad is None:
                    p.grad = gr.clone().to(cuda)
                else:
                    p.grad.copy_(gr.to(cuda))
            if not isinstance(opt, ClampAdam):
                for p in ps:
                    p.grad.clamp_(-5, 5)
            opt.step()

    pa = params()
    a = ClampAdam(pa, lr=1e-2)
    assert a.state_dict()["state"] == {}  # nothing to save before the first step
    run(a, pa, grads[:3])
    sd = a.state_dict()
    assert sorted(sd["state"]) == [0, 1, 2] and float(sd["state"][0]["step"]) == 3.0
    assert sd["state"][2]["exp_avg"].shape == (2, 3, 3) and sd["param_groups"][0]["lr"] == 1e-2
    # resume into a fresh ClampAdam: identical continuation
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    b = ClampAdam(pb, lr=1e-2)
    b.load_state_dict(sd)
    run(a, pa, grads[3:])
    run(b, pb, grads[3:])
    for x, y in zip(pa, pb):
        assert torch.equal(x, y)
    # the same snapshot resumes torch.optim.Adam, and Adam's snapshot resumes ClampAdam
    pc = [torch.nn.Parameter(p.detach().clone()) for p in pb]
    c = torch.optim.Adam(pc, lr=1e-2)
    c.load_state_dict(b.state_dict())
    pd = [torch.nn.Parameter(p.detach().clone()) for p in pb]
    d = ClampAdam(pd, lr=1e-2)
    d.load_state_dict(c.state_dict())
    run(c, pc, grads[:2])
    run(d, pd, grads[:2])
    run(b, pb, grads[:2])
    for x, y, z in zip(pb, pc, pd):
        assert torch.equal(x, z)
        close(y, x, rtol=2e-6, what="torch Adam resumed from a ClampAdam snapshot")


@pytest.mark.parametrize("rows,cols,ld", [(1000, 64, 64), (18560, 4096, 4096), (37, 96, 100), (513, 97, 97), (5, 8, 8)])
def test_colsum(cuda, rows, cols, ld):
    from vistaocr_b200 import ops
    g = torch.Generator().manual_seed(rows + cols)
    x = torch.randn(rows, ld, generator=g)
    out = torch.full((cols,), 7.0, device=cuda)
    ops.colsum(x.to(cuda), rows, cols, ld, out)
    want = x[:, :cols].double().sum(0)
    assert (out.double().cpu() - want).abs().max().item() <= 1e-5 * max(1.0, want.abs().max().item()) * (rows ** 0.5 / 8 + 1)
    ops.colsum(x.to(cuda), rows, cols, ld, out, accumulate=True)
    assert (out.double().cpu() - 2 * want).abs().max().item() <= 2e-5 * max(1.0, want.abs().max().item()) * (rows ** 0.5 / 8 + 1)


def test_first_conv_with_more_than_65535_pixel_tiles(cuda):
    """Large decode batches (cfg5: 512 lines x 30 x 1200 px) give the Cin = 1 convolution > 65535 pixel tiles: the tile
    index must live in grid.x.  Checked on the last image (a convolution is local to its image)."""
    from vistaocr_b200 import ops
    B, H, W, Cout = 40, 30, 7100, 8  # 8.52 M pixels = 66563 tiles of 128
    g = torch.Generator().manual_seed(3)
    x_last = torch.rand((1, 1, H, W), generator=g)
    w = torch.randn((Cout, 1, 3, 3), generator=g) * 0.3
    b = torch.randn((Cout,), generator=g)
    x = torch.zeros((B, H, W, 1), device=cuda)
    x[B - 1] = x_last[0].permute(1, 2, 0).to(cuda)
    z, _ = ops.conv3x3(x, w.to(cuda), b.to(cuda))
    want = F.conv2d(x_last.double(), w.double(), b.double(), padding=1)[0].permute(1, 2, 0)
    got = z[B - 1].double().cpu()
    assert (got - want).abs().max().item() <= 1e-5 * want.abs().max().item()
    assert torch.equal(z[0, 5, 100].cpu(), b)  # an all-zero image yields the bias
