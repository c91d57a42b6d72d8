"""GPU parity of the tcgen05 3xTF32 GEMM (csrc/tc_gemm.cu) against float64, all four operand-major combinations.
Bound: max|C - ref| <= 1e-5 * max|ref| (the fp32 contract); also reported against the FFMA engine."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(A, B, a_mn, b_mn):
    a = A.double().t() if a_mn else A.double()      # -> [M,K]
    b = B.double() if b_mn else B.double().t()      # -> [K,N]
    return a @ b


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (128, 128, 256), (300, 200, 72), (1000, 512, 1024), (64, 96, 2304),
                                   (257, 129, 36)])
def test_tc_gemm_matches_float64(cuda, a_mn, b_mn, M, N, K):
    from vistaocr_b200 import ops
    g = torch.Generator().manual_seed(M + 3 * N + 7 * K + a_mn * 11 + b_mn * 13)
    rup = lambda v: (v + 3) // 4 * 4
    lda = rup(M) if a_mn else rup(K)
    ldb = rup(N) if b_mn else rup(K)
    A = torch.randn((K, lda) if a_mn else (M, lda), generator=g)
    B = torch.randn((K, ldb) if b_mn else (N, ldb), generator=g)
    Av = A[:, :M] if a_mn else A[:, :K]
    Bv = B[:, :N] if b_mn else B[:, :K]
    want = _ref(Av, Bv, a_mn, b_mn)
    bias = torch.randn(N, generator=g)
    C = torch.full((M, N + 1), 3.0, device=cuda)
    As, Bs = ops.split_tf32(A.to(cuda)), ops.split_tf32(B.to(cuda))
    ops.tc_gemm(a_mn, b_mn, M, N, K, As, lda, Bs, ldb, C, N + 1, bias=bias.to(cuda))
    torch.cuda.synchronize()
    got = C[:, :N].double().cpu()
    err = (got - (want + bias.double())).abs().max().item()
    assert err <= 1e-5 * want.abs().max().item(), (err, want.abs().max().item())
    assert (C[:, N] == 3.0).all()
    # accumulate + relu
    C2 = torch.ones((M, N), device=cuda)
    ops.tc_gemm(a_mn, b_mn, M, N, K, As, lda, Bs, ldb, C2, N, relu=True, accumulate=True)
    err2 = (C2.double().cpu() - (want + 1).relu()).abs().max().item()
    assert err2 <= 1e-5 * want.abs().max().item()


def test_split_is_exact(cuda):
    from vistaocr_b200 import ops
    x = torch.randn(100003, device=cuda) * 100
    hi, lo = ops.split_tf32(x)
    assert ((hi + lo) - x).abs().max().item() <= 2.0 ** -21 * x.abs().max().item()  # lo is rounded to TF32 as well
    assert (hi.view(torch.int32) & 0x1fff).abs().max().item() == 0  # hi, lo exactly representable in TF32
    assert (lo.view(torch.int32) & 0x1fff).abs().max().item() == 0


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize("M,N,K,sa,sb", [(128, 128, 64, 1.0, 1.0), (128, 128, 256, 1e-6, 30.0), (300, 200, 72, 1.0, 1.0),
                                         (1000, 512, 1024, 1e4, 1e-7), (64, 96, 2304, 1.0, 1e-3),
                                         (257, 129, 40, 1.0, 1.0), (200, 136, 18816, 1e-5, 1.0)])
def test_tc_gemm_f16_pairs_match_float64(cuda, a_mn, b_mn, M, N, K, sa, sb):
    """kind::f16 on (hi, lo*2^11) planes of the scaled operands: same 1e-5 bound as the TF32 planes, for operand
    magnitudes far outside FP16's own range (the per-tensor power-of-two scale absorbs them)."""
    from vistaocr_b200 import ops
    g = torch.Generator().manual_seed(M + 3 * N + 7 * K + a_mn * 11 + b_mn * 13)
    rup = lambda v: (v + 7) // 8 * 8
    lda = rup(M) if a_mn else rup(K)
    ldb = rup(N) if b_mn else rup(K)
    A = torch.randn((K, lda) if a_mn else (M, lda), generator=g) * sa
    B = torch.randn((K, ldb) if b_mn else (N, ldb), generator=g) * sb
    Av = A[:, :M] if a_mn else A[:, :K]
    Bv = B[:, :N] if b_mn else B[:, :K]
    want = _ref(Av, Bv, a_mn, b_mn)
    scale = want.abs().max().item()
    bias = torch.randn(N, generator=g) * scale
    C = torch.full((M, N + 1), 3.0, device=cuda)
    As, Bs = ops.split_f16(A.to(cuda)), ops.split_f16(B.to(cuda))
    ops.tc_gemm16(a_mn, b_mn, M, N, K, As, lda, Bs, ldb, C, N + 1, bias=bias.to(cuda))
    torch.cuda.synchronize()
    got = C[:, :N].double().cpu()
    err = (got - (want + bias.double())).abs().max().item()
    assert err <= 1e-5 * scale, (err, scale)
    assert (C[:, N] == 3.0).all()
    C2 = torch.full((M, N), scale, device=cuda)
    ops.tc_gemm16(a_mn, b_mn, M, N, K, As, lda, Bs, ldb, C2, N, relu=True, accumulate=True)
    err2 = (C2.double().cpu() - (want + scale).relu()).abs().max().item()
    assert err2 <= 1e-5 * scale


def test_split_f16_pairs(cuda):
    from vistaocr_b200 import ops
    for mag in (1e-30, 1e-8, 1.0, 3e4, 1e20):
        x = torch.randn(100003, device=cuda) * mag
        x[17] = 0.0
        x[18] = x.abs().max() * 2.0 ** -30  # far below the bound: still 11+ bits through the lo plane
        hi, lo, state = ops.split_f16(x)
        e = int(state[0].item())
        back = (hi.double() + lo.double() / 2048.0) * 2.0 ** -e
        m = x.abs().max().item()
        assert 2.0 ** 14 <= m * 2.0 ** e < 2.0 ** 15 or abs(e) == 126  # exponent clamps at +-126
        assert torch.isfinite(hi).all() and torch.isfinite(lo).all()
        rel = ((back - x.double()).abs() / x.double().abs().clamp_min(m * 2.0 ** -28)).max().item()
        assert rel <= 2.0 ** -21, (mag, rel)
        assert abs(back[18].item() - x[18].item()) <= 2.0 ** -10 * abs(x[18].item())
    # caller-supplied bound (no absmax pass): any upper bound works, the scale follows the bound
    x = torch.randn(4096, device=cuda)
    bound = torch.tensor([1000.0], device=cuda)
    hi, lo, state = ops.split_f16(x, bound)
    e = int(state[0].item())
    assert 2.0 ** 14 <= 1000.0 * 2.0 ** e < 2.0 ** 15
    back = (hi.double() + lo.double() / 2048.0) * 2.0 ** -e
    assert (back - x.double()).abs().max().item() <= 2.0 ** -21 * x.abs().max().item()
