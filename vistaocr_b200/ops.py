"""Autograd-aware host wrappers over the C ABI (include/vistaocr_b200.h).

Every op below launches hand-written CUDA kernels through ctypes on torch's current stream; torch itself only
provides device memory, autograd bookkeeping and a few weight-layout views.  There is no fallback path: a missing
library or a non-CUDA tensor raises.
Layouts: CNN activations NHWC fp32 [B,H,W,C]; sequences time-major [T,B,F].
"""
import math

import torch

from . import _lib
from ._lib import check, lib, ptr, stream

F32 = torch.float32


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


def _off(t, elems):
    """Device pointer `elems` fp32 elements into t."""
    import ctypes
    return ctypes.c_void_p(t.data_ptr() + 4 * elems)


# ---------------------------------------------------------------------------------------------------------------
# dense GEMM
# ---------------------------------------------------------------------------------------------------------------
def gemm(transa, transb, M, N, K, A, lda, B, ldb, C, ldc, bias=None, relu=False, accumulate=False, a_off=0, b_off=0,
         c_off=0):
    st = lib().vocr_gemm_f32(int(transa), int(transb), M, N, K, _off(A, a_off), lda, _off(B, b_off), ldb,
                             _off(C, c_off), ldc, ptr(bias), int(relu), int(accumulate), stream())
    check(st, "vocr_gemm_f32")


def colsum(x2d, rows, cols, ld, out, accumulate=False, x_off=0):
    st = lib().vocr_colsum_f32(_off(x2d, x_off), rows, cols, ld, ptr(out), int(accumulate), stream())
    check(st, "vocr_colsum_f32")


def split_tf32(t):
    """(hi, lo) planes of a contiguous fp32 tensor for the 3xTF32 tensor-core GEMM: hi = rna_tf32(t), lo = t - hi."""
    t = _c(t)
    hi = torch.empty_like(t)
    lo = torch.empty_like(t)
    st = lib().vocr_split_tf32_f32(ptr(t), ptr(hi), ptr(lo), t.numel(), stream())
    check(st, "vocr_split_tf32_f32")
    return hi, lo


def _off16(t, elems):
    """Device pointer `elems` fp16 elements into t."""
    import ctypes
    return ctypes.c_void_p(t.data_ptr() + 2 * elems)


def split_f16(t, bound=None):
    """(hi, lo, state) FP16 pair planes of a contiguous fp32 tensor (vocr_split_f16_f32): the planes hold
    t * 2^state[0]; `bound` is an optional device scalar >= max|t| (skips the absmax pass).  Nothing syncs."""
    t = _c(t)
    hi = torch.empty(t.shape, dtype=torch.float16, device=t.device)
    lo = torch.empty(t.shape, dtype=torch.float16, device=t.device)
    state = torch.empty((2,), dtype=torch.int32, device=t.device)
    st = lib().vocr_split_f16_f32(ptr(t), t.numel(), ptr(bound), ptr(state), ptr(hi), ptr(lo), stream())
    check(st, "vocr_split_f16_f32")
    return hi, lo, state


def _splitk_ws(M, N, K, dev):
    tiles = ((M + 127) // 128) * ((N + 127) // 128)
    if K >= 1024 and tiles < 148:  # long reduction, few tiles: allow split-K (the library picks the split count)
        wsb = 4 * M * N * max(1, min(64, (2 * 148) // tiles))
        return torch.empty((wsb,), dtype=torch.uint8, device=dev), wsb
    return None, 0


def tc_gemm16(a_mn, b_mn, M, N, K, A, lda, B, ldb, C, ldc, bias=None, relu=False, accumulate=False, a_off=0, b_off=0,
              c_off=0):
    """Tensor-core GEMM on FP16 pair operands: A = (hi, lo, state), B likewise (see vocr_tc_gemm_f16x3)."""
    ws, wsb = _splitk_ws(M, N, K, C.device)
    st = lib().vocr_tc_gemm_f16x3(int(a_mn), int(b_mn), M, N, K, _off16(A[0], a_off), _off16(A[1], a_off), lda,
                                  ptr(A[2]), _off16(B[0], b_off), _off16(B[1], b_off), ldb, ptr(B[2]), _off(C, c_off),
                                  ldc, ptr(bias), int(relu), int(accumulate), ptr(ws), wsb, _PRODUCTS[0], stream())
    check(st, "vocr_tc_gemm_f16x3")


def tc_gemm(a_mn, b_mn, M, N, K, A, lda, B, ldb, C, ldc, bias=None, relu=False, accumulate=False, a_off=0, b_off=0,
            c_off=0):
    """Tensor-core GEMM on pre-split operands: A = (hi, lo), B = (hi, lo) (see vocr_tc_gemm_tf32x3)."""
    ws, wsb = _splitk_ws(M, N, K, C.device)
    st = lib().vocr_tc_gemm_tf32x3(int(a_mn), int(b_mn), M, N, K, _off(A[0], a_off), _off(A[1], a_off), lda,
                                   _off(B[0], b_off), _off(B[1], b_off), ldb, _off(C, c_off), ldc, ptr(bias),
                                   int(relu), int(accumulate), ptr(ws), wsb, stream())
    check(st, "vocr_tc_gemm_tf32x3")


import os as _os

# Tensor-core (tcgen05 3xTF32) GEMMs are the default; VOCR_TC=0 selects the FFMA engine everywhere (both are this
# library's kernels - this is an engine choice, not a fallback: shapes the TMA path cannot address, i.e. leading
# dimensions that are not multiples of 4 floats, always go to the FFMA engine).
USE_TC = _os.environ.get("VOCR_TC", "1") != "0"
# Operand format of the tensor-core kernels: FP16 pairs (kind::f16, default: twice the K per instruction, half the
# plane bytes) or TF32 planes (VOCR_F16=0, and always where the FP16 alignment rules do not hold: leading dimensions
# that are not multiples of 8, conv channel counts that are not multiples of 64).  Same three-product compensation and
# the same 22 significant bits either way.
USE_F16 = USE_TC and _os.environ.get("VOCR_F16", "1") != "0"


_PRODUCTS = [3]  # arithmetic mode handed to every tensor-core call of this module (per call, not library state)


def set_precision(mode):
    """Arithmetic of the tensor-core GEMM / convolution kernels issued through this module (passed per call):
    "fp32" (default) - three error-compensated products on FP16 pair planes, results within 1e-5 of fp32 (cfg2);
    "fp16"           - one product on the hi planes: fp16 operands (11-bit mantissa, per-tensor power-of-two scaling),
                       fp32 accumulation, fp32 activations / master weights / optimiser - the reduced-precision
                       training mode of BASELINE.json's cfg3.  The BiLSTM recurrence keeps its compensated products.
    Returns the previous mode."""
    prev = get_precision()
    if mode not in ("fp32", "fp16"):
        raise ValueError("precision must be 'fp32' or 'fp16'")
    _PRODUCTS[0] = 1 if mode == "fp16" else 3
    return prev


def get_precision():
    return "fp16" if _PRODUCTS[0] == 1 else "fp32"


class Operand:
    """A GEMM operand: the fp32 tensor plus, lazily, its (hi, lo) TF32 split (shared by every GEMM that reads it)."""

    def __init__(self, t, split=None, split16=None, bound=None):
        self.t = _c(t) if t is not None else None  # None: split-only operand (see _ConvBNReLU.forward)
        self._split = split
        self._split16 = split16
        self.bound = bound  # optional device scalar >= max|t|

    def split(self):
        if self._split is None:
            self._split = split_tf32(self.t)
        return self._split

    def split16(self):
        if self._split16 is None:
            self._split16 = split_f16(self.t, self.bound)
        return self._split16


def mm(transa, transb, M, N, K, A, lda, B, ldb, C, ldc, bias=None, relu=False, accumulate=False, a_off=0, b_off=0,
       c_off=0):
    """C = op(A) op(B) with vocr_gemm_f32 conventions; A, B are `Operand`s.  Picks the tcgen05 kernel when the
    operands are TMA-addressable."""
    ok = USE_TC and lda % 4 == 0 and ldb % 4 == 0 and a_off % 4 == 0 and b_off % 4 == 0 and K >= 1 and \
        A.t.data_ptr() % 16 == 0 and B.t.data_ptr() % 16 == 0
    if ok and USE_F16 and lda % 8 == 0 and ldb % 8 == 0 and a_off % 8 == 0 and b_off % 8 == 0:
        tc_gemm16(transa, 0 if transb else 1, M, N, K, A.split16(), lda, B.split16(), ldb, C, ldc, bias=bias, relu=relu,
                  accumulate=accumulate, a_off=a_off, b_off=b_off, c_off=c_off)
    elif ok:
        tc_gemm(transa, 0 if transb else 1, M, N, K, A.split(), lda, B.split(), ldb, C, ldc, bias=bias, relu=relu,
                accumulate=accumulate, a_off=a_off, b_off=b_off, c_off=c_off)
    else:
        gemm(transa, transb, M, N, K, A.t, lda, B.t, ldb, C, ldc, bias=bias, relu=relu, accumulate=accumulate,
             a_off=a_off, b_off=b_off, c_off=c_off)


class _Linear(torch.autograd.Function):
    """y = x W^T + b (optionally ReLU); x [M,K], W [N,K] (nn.Linear layout)."""

    @staticmethod
    def forward(ctx, x, w, b, relu, x_bound=None):
        x, w = _c(x), _c(w)
        _lib.require_cuda(x, "x", F32)
        M, K = x.shape
        N = w.shape[0]
        y = torch.empty((M, N), dtype=F32, device=x.device)
        xo, wo = Operand(x, bound=x_bound), Operand(w)
        mm(0, 1, M, N, K, xo, K, wo, K, y, N, bias=b, relu=relu)
        ctx.ops = (xo, wo)
        ctx.relu = relu
        ctx.save_for_backward(x, w, y if relu else None)
        ctx.has_bias = b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        dy = _c(dy)
        if ctx.relu:
            dy = torch.where(y > 0, dy, torch.zeros((), dtype=F32, device=dy.device))
        M, K = x.shape
        N = w.shape[0]
        dx = dw = db = None
        xo, wo = ctx.ops
        ctx.ops = None
        if USE_F16 and N % 8 != 0 and K % 8 == 0 and M > 0:
            # an output width that is not a multiple of 8 (alphabet 166 of cfg3) would send both backward GEMMs to the
            # FFMA engine (dy is their reduction- / row-major operand with leading dimension N): pad dy and W with zero
            # columns / rows to the next multiple of 8 instead - two small copies
            Np = (N + 7) // 8 * 8
            dyp = torch.zeros((M, Np), dtype=F32, device=dy.device)
            dyp[:, :N].copy_(dy)
            dyo = Operand(dyp)
            if ctx.needs_input_grad[0]:
                wp = torch.zeros((Np, K), dtype=F32, device=dy.device)
                wp[:N].copy_(w)
                dx = torch.empty_like(x)
                mm(0, 0, M, K, Np, dyo, Np, Operand(wp), K, dx, K)
            if ctx.needs_input_grad[1]:
                dwp = torch.empty((Np, K), dtype=F32, device=dy.device)
                mm(1, 0, Np, K, M, dyo, Np, xo, K, dwp, K)
                dw = dwp[:N]
        else:
            dyo = Operand(dy)
            if ctx.needs_input_grad[0]:
                dx = torch.empty_like(x)
                mm(0, 0, M, K, N, dyo, N, wo, K, dx, K)
            if ctx.needs_input_grad[1]:
                dw = torch.empty_like(w)
                mm(1, 0, N, K, M, dyo, N, xo, K, dw, K)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = torch.empty((N,), dtype=F32, device=x.device)
            colsum(dy, M, N, N, db)
        return dx, dw, db, None, None


def linear(x, w, b=None, relu=False, x_bound=None):
    """x_bound: optional device scalar >= max|x| (spares the operand split its abs-max pass)."""
    return _Linear.apply(x, w, b, relu, x_bound)


def linear_argmax_supported(K, N):
    """Shapes vocr_tc_gemm_f16x3_argmax serves (one output tile per row block, the persistent short-K kernel)."""
    return USE_F16 and N <= 128 and K % 8 == 0 and K <= 1536


def linear_argmax(x, w, b, lens_dev, T, B, thresh, x_bound=None):
    """Inference only: frame labels path[B,T] (int32; see vocr_greedy_decode_f32) of the logits x W^T + b, x [T*B,K] -
    the arg-max runs in the GEMM epilogue and the logits are never written."""
    x, w = _c(x), _c(w.detach())
    _lib.require_cuda(x, "x", F32)
    M, K = x.shape
    N = w.shape[0]
    if M != T * B or not linear_argmax_supported(K, N):
        raise _lib.VocrError("linear_argmax: unsupported shape")
    xs = Operand(x, bound=x_bound).split16()
    ws = Operand(w).split16()
    path = torch.empty((B, max(T, 1)), dtype=torch.int32, device=x.device)
    import numpy as _np
    st = lib().vocr_tc_gemm_f16x3_argmax(M, N, K, ptr(xs[0]), ptr(xs[1]), K, ptr(xs[2]), ptr(ws[0]), ptr(ws[1]), K,
                                         ptr(ws[2]), None, N, ptr(b), ptr(lens_dev), T, B, float(_np.float32(thresh)),
                                         ptr(path), _PRODUCTS[0], stream())
    check(st, "vocr_tc_gemm_f16x3_argmax")
    return path


# ---------------------------------------------------------------------------------------------------------------
# conv3x3 + BatchNorm + ReLU block
# ---------------------------------------------------------------------------------------------------------------
def _conv_fwd(x, wk, bias, B, H, W, Cin, Cout, stats, zmax=None):
    z = torch.empty((B, H, W, Cout), dtype=F32, device=x.device)
    st = lib().vocr_conv3x3_fwd_f32(ptr(x), ptr(wk), ptr(bias), ptr(z), B, H, W, Cin, Cout, ptr(stats), ptr(zmax),
                                    stream())
    check(st, "vocr_conv3x3_fwd_f32")
    return z


def _weight_layout(w, want_k, want_d):
    Cout, Cin = w.shape[0], w.shape[1]
    wk = torch.empty((9 * Cin, Cout), dtype=F32, device=w.device) if want_k else None
    wd = torch.empty((9 * Cout, Cin), dtype=F32, device=w.device) if want_d else None
    st = lib().vocr_conv_weight_layout_f32(ptr(w), Cin, Cout, ptr(wk), ptr(wd), stream())
    check(st, "vocr_conv_weight_layout_f32")
    return wk, wd


def _conv_wgrad(x, dz, B, H, W, Cin, Cout):
    dw = torch.empty((Cout, Cin, 3, 3), dtype=F32, device=x.device)
    wsb = lib().vocr_conv3x3_wgrad_workspace_size(B, H, W, Cin, Cout)
    ws = torch.empty((wsb,), dtype=torch.uint8, device=x.device)
    st = lib().vocr_conv3x3_wgrad_f32(ptr(x), ptr(dz), ptr(dw), B, H, W, Cin, Cout, ptr(ws), wsb, stream())
    check(st, "vocr_conv3x3_wgrad_f32")
    return dw


def _tc_conv_fwd(x_s, w_s, bias, B, H, W, Cin, Cout):
    """x_s, w_s: (hi, lo) planes; x NHWC [B,H,W,Cin], w K-major [Cout, 9*Cin]."""
    z = torch.empty((B, H, W, Cout), dtype=F32, device=x_s[0].device)
    st = lib().vocr_tc_conv3x3_fwd(ptr(x_s[0]), ptr(x_s[1]), ptr(w_s[0]), ptr(w_s[1]), ptr(bias), ptr(z), B, H, W, Cin,
                                   Cout, stream())
    check(st, "vocr_tc_conv3x3_fwd")
    return z


def _tc_conv_wgrad(x_s, dz_s, B, H, W, Cin, Cout):
    dev = x_s[0].device
    dw = torch.empty((Cout, Cin, 3, 3), dtype=F32, device=dev)
    wsb = lib().vocr_tc_conv3x3_wgrad_workspace_size(B, H, W, Cin, Cout)
    ws = torch.empty((wsb,), dtype=torch.uint8, device=dev)
    st = lib().vocr_tc_conv3x3_wgrad(ptr(x_s[0]), ptr(x_s[1]), ptr(dz_s[0]), ptr(dz_s[1]), ptr(dw), B, H, W, Cin, Cout,
                                     ptr(ws), wsb, stream())
    check(st, "vocr_tc_conv3x3_wgrad")
    return dw


def _tc_conv_fwd16(x_s, w_s, bias, B, H, W, Cin, Cout, zmax=None):
    """x_s, w_s: (hi, lo, state) FP16 pair planes; x NHWC [B,H,W,Cin], w K-major [Cout, 9*Cin].  zmax: optional zeroed
    device scalar that receives max |z| from the epilogue."""
    z = torch.empty((B, H, W, Cout), dtype=F32, device=x_s[0].device)
    st = lib().vocr_tc_conv3x3_fwd_f16(ptr(x_s[0]), ptr(x_s[1]), ptr(x_s[2]), ptr(w_s[0]), ptr(w_s[1]), ptr(w_s[2]),
                                       ptr(bias), ptr(z), B, H, W, Cin, Cout, _PRODUCTS[0], ptr(zmax), stream())
    check(st, "vocr_tc_conv3x3_fwd_f16")
    return z


def _tc_conv_wgrad16(x_s, dz_s, B, H, W, Cin, Cout):
    dev = x_s[0].device
    dw = torch.empty((Cout, Cin, 3, 3), dtype=F32, device=dev)
    wsb = lib().vocr_tc_conv3x3_wgrad_workspace_size(B, H, W, Cin, Cout)
    ws = torch.empty((wsb,), dtype=torch.uint8, device=dev)
    st = lib().vocr_tc_conv3x3_wgrad_f16(ptr(x_s[0]), ptr(x_s[1]), ptr(x_s[2]), ptr(dz_s[0]), ptr(dz_s[1]),
                                         ptr(dz_s[2]), ptr(dw), B, H, W, Cin, Cout, ptr(ws), wsb, _PRODUCTS[0], stream())
    check(st, "vocr_tc_conv3x3_wgrad_f16")
    return dw


def im2col3x3_f16(x, bound=None):
    """FP16 pair planes [B*H*W, 9*C] of the 3x3 / pad 1 patches of a narrow NHWC activation (vocr_im2col3x3_f16)."""
    x = _c(x)
    B, H, W, C = x.shape
    hi = torch.empty((B * H * W, 9 * C), dtype=torch.float16, device=x.device)
    lo = torch.empty((B * H * W, 9 * C), dtype=torch.float16, device=x.device)
    state = torch.empty((2,), dtype=torch.int32, device=x.device)
    st = lib().vocr_im2col3x3_f16(ptr(x), B, H, W, C, ptr(bound), ptr(state), ptr(hi), ptr(lo), stream())
    check(st, "vocr_im2col3x3_f16")
    return hi, lo, state


def _narrow(Cin, Cout):
    """Conv layers with too few input channels for the implicit-GEMM kernels run as K = 9*Cin GEMMs over patch planes."""
    return USE_F16 and Cin < 64 and Cin % 8 == 0 and Cout % 8 == 0


def conv3x3(x, weight, bias, x_op=None, zmax=None):
    """z = conv3x3_pad1(x) + bias, NHWC; picks the tensor-core kernel when Cin % 32 == 0.  Returns (z, x_operand).
    zmax (optional zeroed device scalar) receives max |z| when the FP16-pair implicit-GEMM kernel runs (it stays 0 on the
    other paths: callers must treat 0 as "not measured")."""
    B, H, W, Cin = x.shape
    Cout = weight.shape[0]
    if _narrow(Cin, Cout):
        if x_op is None:
            x_op = Operand(None, split16=im2col3x3_f16(x, getattr(x, "_vocr_bound", None)))
            x_op.cols = True
        wn = Operand(weight.detach().permute(0, 2, 3, 1).reshape(Cout, 9 * Cin))
        z = torch.empty((B, H, W, Cout), dtype=F32, device=x.device)
        tc_gemm16(0, 0, B * H * W, Cout, 9 * Cin, x_op.split16(), 9 * Cin, wn.split16(), 9 * Cin, z, Cout, bias=bias)
        return z, x_op
    if USE_F16 and Cin % 64 == 0 and Cout % 4 == 0:
        x_op = x_op or getattr(x, "_vocr_op", None) or Operand(x, bound=getattr(x, "_vocr_bound", None))
        wn = Operand(weight.detach().permute(0, 2, 3, 1).reshape(Cout, 9 * Cin))
        if zmax is not None:
            zmax.measured = True
        return _tc_conv_fwd16(x_op.split16(), wn.split16(), bias, B, H, W, Cin, Cout, zmax), x_op
    if USE_TC and Cin % 32 == 0 and Cout % 4 == 0:
        x_op = x_op or getattr(x, "_vocr_op", None) or Operand(x)
        wn = Operand(weight.detach().permute(0, 2, 3, 1).reshape(Cout, 9 * Cin))
        return _tc_conv_fwd(x_op.split(), wn.split(), bias, B, H, W, Cin, Cout), x_op
    wk, _ = _weight_layout(_c(weight.detach()), True, False)
    if zmax is not None:
        zmax.measured = True
    return _conv_fwd(x, wk, bias, B, H, W, Cin, Cout, None, zmax), x_op


def conv3x3_dgrad(dz, weight, dz_op=None):
    """dx = conv3x3 data gradient (NHWC)."""
    B, H, W, Cout = dz.shape
    Cin = weight.shape[1]
    if USE_F16 and Cout % 64 == 0 and Cin % 4 == 0:
        dz_op = dz_op or Operand(dz)
        wd = Operand(weight.detach().flip(2, 3).permute(1, 2, 3, 0).reshape(Cin, 9 * Cout))
        return _tc_conv_fwd16(dz_op.split16(), wd.split16(), None, B, H, W, Cout, Cin), dz_op
    if USE_TC and Cout % 32 == 0 and Cin % 4 == 0:
        dz_op = dz_op or Operand(dz)
        wd = Operand(weight.detach().flip(2, 3).permute(1, 2, 3, 0).reshape(Cin, 9 * Cout))
        return _tc_conv_fwd(dz_op.split(), wd.split(), None, B, H, W, Cout, Cin), dz_op
    _, wd = _weight_layout(_c(weight.detach()), False, True)
    return _conv_fwd(dz, wd, None, B, H, W, Cout, Cin, None), dz_op


def conv3x3_wgrad(x, dz, x_op=None, dz_op=None):
    B, H, W, Cin = x.shape
    Cout = dz.shape[3]
    if _narrow(Cin, Cout):
        if x_op is None or not getattr(x_op, "cols", False):
            x_op = Operand(None, split16=im2col3x3_f16(x))
        dz_op = dz_op or Operand(dz)
        dwt = torch.empty((9 * Cin, Cout), dtype=F32, device=x.device)  # [(ky, kx, ci), co] = cols^T dz
        tc_gemm16(1, 1, 9 * Cin, Cout, B * H * W, x_op.split16(), 9 * Cin, dz_op.split16(), Cout, dwt, Cout)
        return dwt.view(3, 3, Cin, Cout).permute(3, 2, 0, 1).contiguous()
    if USE_F16 and Cin % 64 == 0 and Cout % 64 == 0:
        x_op = x_op or Operand(x)
        dz_op = dz_op or Operand(dz)
        return _tc_conv_wgrad16(x_op.split16(), dz_op.split16(), B, H, W, Cin, Cout)
    if USE_TC and Cin % 32 == 0 and Cout % 32 == 0:
        if x_op is None or (x_op.t is None and x_op._split is None):
            x_op = Operand(x)  # a split-only operand that carries FP16 planes only: rebuild the TF32 planes from x
        dz_op = dz_op or Operand(dz)
        return _tc_conv_wgrad(x_op.split(), dz_op.split(), B, H, W, Cin, Cout)
    return _conv_wgrad(x, dz, B, H, W, Cin, Cout)


def colstats(z, C, stats):
    st = lib().vocr_colstats_f32(ptr(z), z.numel() // C, C, ptr(stats), stream())
    check(st, "vocr_colstats_f32")


# Inference: Conv + BatchNorm (running statistics) + ReLU as ONE kernel (vocr_tc_conv3x3_bnrelu_f16); VOCR_FUSE_EVAL=0
# keeps the conv -> z -> bn_relu_apply sequence (same values in the fp32 activation, bit for bit).
FUSE_EVAL = _os.environ.get("VOCR_FUSE_EVAL", "1") != "0"


def _conv_bn_relu_eval_fused(x, weight, bias, gamma, beta, running_mean, running_var, eps, seq_layout, planes,
                             allow_planes_only):
    """a = relu(bn_eval(conv3x3(x) + bias)) without the fp32 z round trip: scale / shift are known before the convolution
    runs, an upper bound of |a| follows from the weights and the bound of x (vocr_bn_eval_bound_f32), so the conv epilogue
    writes the activation and / or the next convolution's FP16 pair planes itself."""
    B, H, W, Cin = x.shape
    Cout = weight.shape[0]
    dev = x.device
    x_op = getattr(x, "_vocr_op", None) or Operand(x, bound=getattr(x, "_vocr_bound", None))
    xs = x_op.split16()
    xb = getattr(x, "_vocr_bound", None)
    if xb is None:
        xb = x_op.bound
    # no bound known: the split's abs-max pass left max |x| (as float bits) in state[1]
    xb_ptr = ptr(xb) if xb is not None else _off(xs[2], 1)
    scale = torch.empty((Cout,), dtype=F32, device=dev)
    shift = torch.empty((Cout,), dtype=F32, device=dev)
    aux = torch.empty((2,), dtype=F32, device=dev)  # [0] activation bound, [1] max|scale|
    st = lib().vocr_bn_finalize_f32(None, B * H * W, ptr(gamma), ptr(beta), ptr(running_mean), ptr(running_var), 0.0,
                                    float(eps), 0, ptr(scale), ptr(shift), None, None, Cout, ptr(aux), None, stream())
    check(st, "vocr_bn_finalize_f32")
    w = _c(weight.detach())
    st = lib().vocr_bn_eval_bound_f32(ptr(w), ptr(bias), ptr(scale), ptr(shift), xb_ptr, Cout, 9 * Cin, ptr(aux),
                                      stream())
    check(st, "vocr_bn_eval_bound_f32")
    wn = Operand(w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin))
    ws = wn.split16()
    if seq_layout:
        a = torch.empty((W, B, H * Cout), dtype=F32, device=dev)
        strides = (H * Cout, Cout, B * H * Cout)
    else:
        a = torch.empty((B, H, W, Cout), dtype=F32, device=dev)  # (not written in the planes-only case)
        strides = (H * W * Cout, W * Cout, Cout)
    want16 = not seq_layout and Cout % 64 == 0 and planes
    planes_only = allow_planes_only and want16 and not x.requires_grad
    a_hi16 = torch.empty((B, H, W, Cout), dtype=torch.float16, device=dev) if want16 else None
    a_lo16 = torch.empty((B, H, W, Cout), dtype=torch.float16, device=dev) if want16 else None
    pstate = torch.empty((2,), dtype=torch.int32, device=dev) if want16 else None
    amax = torch.zeros((1,), dtype=F32, device=dev)  # measured max a: where the next block's analytic bound starts
    st = lib().vocr_tc_conv3x3_bnrelu_f16(ptr(xs[0]), ptr(xs[1]), ptr(xs[2]), ptr(ws[0]), ptr(ws[1]), ptr(ws[2]),
                                          ptr(bias), ptr(scale), ptr(shift), None if planes_only else ptr(a),
                                          strides[0], strides[1], strides[2], ptr(a_hi16), ptr(a_lo16),
                                          ptr(aux) if want16 else None, ptr(pstate), ptr(amax), B, H, W, Cin, Cout,
                                          _PRODUCTS[0], stream())
    check(st, "vocr_tc_conv3x3_bnrelu_f16")
    if want16:
        a._vocr_op = Operand(None, None, (a_hi16, a_lo16, pstate))
    a._vocr_bound = amax  # (the planes above are scaled for the analytic bound aux[0] >= amax)
    a._vocr_plane_bound = aux
    return a


class _ConvBNReLU(torch.autograd.Function):
    """a = relu(batchnorm(conv3x3(x) + bias)).  x NHWC [B,H,W,Cin]; weight [Cout,Cin,3,3] (state_dict layout).
    seq_layout=True writes a as the time-major sequence [W, B, H*Cout] (feature = y*Cout + c)."""

    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, running_mean, running_var, training, momentum, eps, seq_layout,
                planes=True, allow_planes_only=False):
        x = _c(x)
        _lib.require_cuda(x, "x", F32)
        B, H, W, Cin = x.shape
        Cout = weight.shape[0]
        dev = x.device
        stats = torch.zeros((2 * Cout,), dtype=torch.float64, device=dev) if training else None
        # eval mode: the convolution's epilogue measures max |z|, from which bn_finalize bounds the activations - the
        # apply kernel can then emit the next layer's FP16 pair planes directly (no abs-max + split passes)
        zmax = torch.zeros((1,), dtype=F32, device=dev) if (USE_F16 and not training) else None
        z, x_op = conv3x3(x, weight, bias, zmax=zmax)
        if zmax is not None and not getattr(zmax, "measured", False):
            zmax = None
        if training:
            colstats(z, Cout, stats)
        ctx.x_op = x_op
        scale = torch.empty((Cout,), dtype=F32, device=dev)
        shift = torch.empty((Cout,), dtype=F32, device=dev)
        mean = torch.empty((Cout,), dtype=F32, device=dev)
        invstd = torch.empty((Cout,), dtype=F32, device=dev)
        aux = torch.empty((2,), dtype=F32, device=dev)  # [0] activation bound, [1] max|scale| (FP16 pair planes)
        st = lib().vocr_bn_finalize_f32(ptr(stats), B * H * W, ptr(gamma), ptr(beta), ptr(running_mean),
                                        ptr(running_var), float(momentum), float(eps), int(training), ptr(scale),
                                        ptr(shift), ptr(mean), ptr(invstd), Cout, ptr(aux), ptr(zmax), stream())
        check(st, "vocr_bn_finalize_f32")
        if seq_layout:
            a = torch.empty((W, B, H * Cout), dtype=F32, device=dev)
            strides = (H * Cout, Cout, B * H * Cout)
        else:
            a = torch.empty((B, H, W, Cout), dtype=F32, device=dev)  # (not written in the planes-only case below)
            strides = (H * W * Cout, W * Cout, Cout)
        # the next tensor-core conv reads this activation as TF32 (hi, lo) planes: let the apply kernel write them
        want_split = USE_TC and not USE_F16 and not seq_layout and Cout % 32 == 0 and planes
        a_hi = torch.empty_like(a) if want_split else None
        a_lo = torch.empty_like(a) if want_split else None
        # ... or as FP16 pair planes (the analytic bound of bn_finalize needs batch statistics)
        bounded = training or zmax is not None  # aux[0] holds a valid bound of the activations
        want16 = USE_F16 and bounded and not seq_layout and Cout % 64 == 0 and planes
        a_hi16 = torch.empty(a.shape, dtype=torch.float16, device=dev) if want16 else None
        a_lo16 = torch.empty(a.shape, dtype=torch.float16, device=dev) if want16 else None
        pstate = torch.empty((2,), dtype=torch.int32, device=dev) if want16 else None
        # inference with a tensor-core conv as the only consumer (`planes`): that conv reads the planes, nobody reads the
        # fp32 activation - do not write it (the returned tensor then only carries the shape and the operand planes)
        # (training too: the consumer's backward multiplies by the planes as well - conv3x3_wgrad's FP16 path)
        planes_only = allow_planes_only and want16
        st = lib().vocr_bn_relu_apply_f32(ptr(z), ptr(scale), ptr(shift), None if planes_only else ptr(a), ptr(a_hi),
                                          ptr(a_lo), B, H, W, Cout,
                                          strides[0], strides[1], strides[2], ptr(a_hi16), ptr(a_lo16),
                                          ptr(aux) if want16 else None, ptr(pstate), stream())
        check(st, "vocr_bn_relu_apply_f32")
        if want_split or want16:
            # split-only: an Operand that held `a` would form a reference cycle (a -> attribute -> a) and keep the
            # activation and both planes alive until Python's cyclic GC runs - several GB per step
            a._vocr_op = Operand(None, (a_hi, a_lo) if want_split else None,
                                 (a_hi16, a_lo16, pstate) if want16 else None)
        if USE_F16 and bounded:
            a._vocr_bound = aux  # aux[0] bounds a (and anything pooled from it): lets a later split skip its absmax
        ctx.save_for_backward(x, weight, z, scale, shift, mean, invstd, aux)
        ctx.dims = (B, H, W, Cin, Cout, strides, bool(training))
        return a

    @staticmethod
    def backward(ctx, da):
        x, weight, z, scale, shift, mean, invstd, aux = ctx.saved_tensors
        B, H, W, Cin, Cout, strides, training = ctx.dims
        dev = x.device
        da = _c(da)
        dz = torch.empty((B, H, W, Cout), dtype=F32, device=dev)  # (virtual until written: see planes_only below)
        dgamma = torch.empty((Cout,), dtype=F32, device=dev)
        dbeta = torch.empty((Cout,), dtype=F32, device=dev)
        dbias = torch.empty((Cout,), dtype=F32, device=dev)
        red = torch.empty((3 * Cout,), dtype=torch.float64, device=dev)
        want_split = USE_TC and not USE_F16 and Cout % 32 == 0 and (Cin % 32 == 0 or ctx.needs_input_grad[0])
        dz_hi = torch.empty_like(dz) if want_split else None
        dz_lo = torch.empty_like(dz) if want_split else None
        want16 = USE_F16 and Cout % 64 == 0 and (Cin % 64 == 0 or ctx.needs_input_grad[0])
        dz_hi16 = torch.empty(dz.shape, dtype=torch.float16, device=dev) if want16 else None
        dz_lo16 = torch.empty(dz.shape, dtype=torch.float16, device=dev) if want16 else None
        pstate = torch.empty((2,), dtype=torch.int32, device=dev) if want16 else None
        # both gradient convolutions read the FP16 pair planes when their shapes allow it: the fp32 dz is then never read
        # and not written (a third of this pass's stores)
        planes_only = want16 and (Cin % 64 == 0 or _narrow(Cin, Cout)) and (Cin % 4 == 0 or not ctx.needs_input_grad[0])
        st = lib().vocr_bn_relu_bwd_f32(ptr(da), ptr(z), ptr(scale), ptr(shift), ptr(mean), ptr(invstd),
                                        int(training), B, H, W, Cout, strides[0], strides[1], strides[2],
                                        None if planes_only else ptr(dz), ptr(dz_hi), ptr(dz_lo), ptr(dgamma), ptr(dbeta), ptr(dbias), ptr(red),
                                        ptr(dz_hi16), ptr(dz_lo16), _off(aux, 1) if want16 else None, ptr(pstate),
                                        stream())
        check(st, "vocr_bn_relu_bwd_f32")
        dx = None
        dz_op = None
        if want_split or want16:
            dz_op = Operand(dz, (dz_hi, dz_lo) if want_split else None,
                            (dz_hi16, dz_lo16, pstate) if want16 else None)
        if ctx.needs_input_grad[0]:
            dx, dz_op = conv3x3_dgrad(dz, weight, dz_op)
        dw = conv3x3_wgrad(x, dz, ctx.x_op, dz_op)
        ctx.x_op = None
        return dx, dw, dbias, dgamma, dbeta, None, None, None, None, None, None, None, None


def conv_bn_relu(x, weight, bias, gamma, beta, running_mean, running_var, training, momentum=0.1, eps=1e-5,
                 seq_layout=False, planes=True, allow_planes_only=False):
    """planes: the consumer is another tensor-core conv, so the apply kernel also emits the operand planes (pass False
    when a pooling layer follows - the planes would go unread).  allow_planes_only: the caller guarantees that a
    tensor-core conv of this module (Cin % 64 == 0: forward, weight gradient and data gradient all read the FP16 pair
    planes) is the ONLY consumer; the fp32 activation is then not written at all - the returned tensor carries the shape
    and the planes, its values are undefined."""
    Cin, Cout = x.shape[3], weight.shape[0]
    no_grad = not torch.is_grad_enabled() or not any(
        t is not None and t.requires_grad for t in (x, weight, bias, gamma, beta))
    if FUSE_EVAL and USE_F16 and not training and no_grad and Cin % 64 == 0 and Cout % 4 == 0 and \
            not _narrow(Cin, Cout) and x.is_cuda and x.dtype == F32:
        return _conv_bn_relu_eval_fused(_c(x), weight, bias, gamma, beta, running_mean, running_var, eps, seq_layout,
                                        planes, allow_planes_only)
    return _ConvBNReLU.apply(x, weight, bias, gamma, beta, running_mean, running_var, training, momentum, eps,
                             seq_layout, planes, allow_planes_only)


class _RapidDS(torch.autograd.Function):
    """y = maxpool2x2(relu(conv3x3(x) + bias)), Cout = 16 (reference cnnlstm.py:114-121)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x = _c(x)
        _lib.require_cuda(x, "x", F32)
        B, H, W, Cin = x.shape
        assert weight.shape[0] == 16
        dev = x.device
        wk, _ = _weight_layout(_c(weight), True, False)
        y = torch.empty((B, H // 2, W // 2, 16), dtype=F32, device=dev)
        need_bwd = any(ctx.needs_input_grad)
        arg = torch.empty((B, H // 2, W // 2, 16), dtype=torch.uint8, device=dev) if need_bwd else None
        st = lib().vocr_rds_fwd_f32(ptr(x), ptr(wk), ptr(bias), ptr(y), ptr(arg), B, H, W, Cin, stream())
        check(st, "vocr_rds_fwd_f32")
        ctx.save_for_backward(x, weight, y, arg)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, y, arg = ctx.saved_tensors
        B, H, W, Cin = x.shape
        dev = x.device
        dy = _c(dy)
        if Cin == 1 and not ctx.needs_input_grad[0]:
            dw = torch.empty((16, 1, 3, 3), dtype=F32, device=dev)
            db = torch.empty((16,), dtype=F32, device=dev)
            ws = torch.empty((160,), dtype=torch.float64, device=dev)
            st = lib().vocr_rds_wgrad_c1_f32(ptr(x), ptr(dy), ptr(y), ptr(arg), ptr(dw), ptr(db), B, H, W, ptr(ws),
                                             stream())
            check(st, "vocr_rds_wgrad_c1_f32")
            return None, dw, db
        dpre = torch.empty((B, H, W, 16), dtype=F32, device=dev)
        st = lib().vocr_rds_unpool_f32(ptr(dy), ptr(y), ptr(arg), ptr(dpre), B, H, W, stream())
        check(st, "vocr_rds_unpool_f32")
        if Cin == 16:  # second stage (line height 120): direct 16 -> 16 kernels
            dx = None
            if ctx.needs_input_grad[0]:
                _, wd = _weight_layout(_c(weight), False, True)
                dx = torch.empty((B, H, W, 16), dtype=F32, device=dev)
                st = lib().vocr_conv3x3_c16_fwd_f32(ptr(dpre), ptr(wd), None, ptr(dx), B, H, W, stream())
                check(st, "vocr_conv3x3_c16_fwd_f32")
            dw = torch.empty((16, 16, 3, 3), dtype=F32, device=dev)
            db = torch.empty((16,), dtype=F32, device=dev)
            ws = torch.empty((2320,), dtype=torch.float64, device=dev)
            st = lib().vocr_conv3x3_c16_wgrad_f32(ptr(x), ptr(dpre), ptr(dw), ptr(db), B, H, W, ptr(ws), stream())
            check(st, "vocr_conv3x3_c16_wgrad_f32")
            return dx, dw, db
        dx = None
        if ctx.needs_input_grad[0]:
            _, wd = _weight_layout(_c(weight), False, True)
            dx = _conv_fwd(dpre, wd, None, B, H, W, 16, Cin, None)
        dw = _conv_wgrad(x, dpre, B, H, W, Cin, 16)
        db = torch.empty((16,), dtype=F32, device=dev)
        colsum(dpre, B * H * W, 16, 16, db)
        return dx, dw, db


def rapid_ds(x, weight, bias):
    return _RapidDS.apply(x, weight, bias)


class _FracPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, samples):
        x = _c(x)
        _lib.require_cuda(x, "x", F32)
        B, H, W, C = x.shape
        Ho, Wo = int(H * 0.5), int(W * 0.7)  # same float64 product + truncation as ATen / the reference
        samples = _c(samples.to(device=x.device, dtype=F32))
        assert samples.shape == (B, C, 2)
        y = torch.empty((B, Ho, Wo, C), dtype=F32, device=x.device)
        idx = torch.empty((B, Ho, Wo, C), dtype=torch.int32, device=x.device) if x.requires_grad else None
        st = lib().vocr_fracpool_fwd_f32(ptr(x), ptr(samples), ptr(y), ptr(idx), B, H, W, C, Ho, Wo, stream())
        check(st, "vocr_fracpool_fwd_f32")
        ctx.save_for_backward(idx)
        ctx.dims = (B, H, W, C, Ho, Wo)
        return y

    @staticmethod
    def backward(ctx, dy):
        (idx,) = ctx.saved_tensors
        B, H, W, C, Ho, Wo = ctx.dims
        dy = _c(dy)
        dx = torch.empty((B, H, W, C), dtype=F32, device=dy.device)
        st = lib().vocr_fracpool_bwd_f32(ptr(dy), ptr(idx), ptr(dx), B, H, W, C, Ho, Wo, stream())
        check(st, "vocr_fracpool_bwd_f32")
        return dx, None


def fracpool(x, samples):
    y = _FracPool.apply(x, samples)
    bound = getattr(x, "_vocr_bound", None)
    if bound is not None:
        y._vocr_bound = bound  # a max-pool cannot exceed its input
    return y


# ---------------------------------------------------------------------------------------------------------------
# bidirectional LSTM layer
# ---------------------------------------------------------------------------------------------------------------
class _BiLSTMLayer(torch.autograd.Function):
    """x [T,B,Din]; w_ih [8H,Din] (forward rows then reverse rows); w_hh [2,4H,H]; bias [8H] (= b_ih + b_hh);
    lens_dev int32 [B] on the device; tmax = max(lens) (python int).  Returns out [T,B,2H]."""

    @staticmethod
    def forward(ctx, x, w_ih, w_hh, bias, lens_dev, tmax, save, x_bound=None):
        x, w_ih, w_hh, bias = _c(x), _c(w_ih), _c(w_hh), _c(bias)
        _lib.require_cuda(x, "x", F32)
        T, B, Din = x.shape
        H = w_hh.shape[2]
        dev = x.device
        xproj = torch.empty((T, B, 2, 4 * H), dtype=F32, device=dev)
        xo, wo = Operand(x, bound=x_bound), Operand(w_ih)
        mm(0, 1, T * B, 8 * H, Din, xo, Din, wo, Din, xproj, 8 * H, bias=bias)
        out = torch.empty((T, B, 2 * H), dtype=F32, device=dev)
        gates = torch.empty((T, B, 2, 4 * H), dtype=F32, device=dev) if save else None
        cst = torch.empty((T, B, 2, H), dtype=F32, device=dev) if save else None
        wsb = lib().vocr_bilstm_workspace_size(int(tmax), B, H, 0)
        if wsb == 0:
            raise _lib.VocrError("vocr_bilstm: unsupported hidden size %d (max 512)" % H)
        ws = torch.empty((wsb,), dtype=torch.uint8, device=dev)
        st = lib().vocr_bilstm_fwd_f32(ptr(xproj), ptr(w_hh), ptr(lens_dev), ptr(out), ptr(gates), ptr(cst), T, B, H,
                                       int(tmax), ptr(ws), wsb, stream())
        check(st, "vocr_bilstm_fwd_f32")
        if save:
            ctx.save_for_backward(x, w_ih, w_hh, lens_dev, gates, cst, out)
            ctx.tmax = int(tmax)
            ctx.ops = (xo, wo)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, w_ih, w_hh, lens_dev, gates, cst, out = ctx.saved_tensors
        T, B, Din = x.shape
        H = w_hh.shape[2]
        dev = x.device
        dout = _c(dout)
        dgates = torch.empty((T, B, 2, 4 * H), dtype=F32, device=dev)
        wsb = lib().vocr_bilstm_workspace_size(ctx.tmax, B, H, 1)
        ws = torch.empty((wsb,), dtype=torch.uint8, device=dev)
        dg_max = torch.empty((1,), dtype=F32, device=dev)  # max |dgates| from the kernel: no abs-max pass for the GEMMs
        st = lib().vocr_bilstm_bwd_f32(ptr(dout), ptr(w_hh), ptr(lens_dev), ptr(gates), ptr(cst), ptr(dgates), ptr(dg_max),
                                       T, B, H, ctx.tmax, ptr(ws), wsb, stream())
        check(st, "vocr_bilstm_bwd_f32")
        dx = dw_ih = dw_hh = db = None
        M = T * B
        xo, wo = ctx.ops
        ctx.ops = None
        dgo, outo = Operand(dgates, bound=dg_max), Operand(out, bound=const_scalar(dev, 1.0))  # |h| < 1
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            mm(0, 0, M, Din, 8 * H, dgo, 8 * H, wo, Din, dx, Din)
        if ctx.needs_input_grad[1]:
            dw_ih = torch.empty_like(w_ih)
            mm(1, 0, 8 * H, Din, M, dgo, 8 * H, xo, Din, dw_ih, Din)
        if ctx.needs_input_grad[2]:
            # dW_hh[d] = sum_t dgates[t,:,d,:]^T h_prev[t,:,d,:]; h_prev is `out` shifted one step along the
            # direction's time order (rows beyond a sample's length are zero in both operands)
            dw_hh = torch.zeros_like(w_hh)
            if T > 1:
                K = (T - 1) * B
                mm(1, 0, 4 * H, H, K, dgo, 8 * H, outo, 2 * H, dw_hh, H, a_off=B * 8 * H, b_off=0, c_off=0)
                mm(1, 0, 4 * H, H, K, dgo, 8 * H, outo, 2 * H, dw_hh, H, a_off=4 * H, b_off=B * 2 * H + H,
                   c_off=4 * H * H)
        if ctx.needs_input_grad[3]:
            db = torch.empty((8 * H,), dtype=F32, device=dev)
            colsum(dgates, M, 8 * H, 8 * H, db)
        return dx, dw_ih, dw_hh, db, None, None, None, None


def bilstm_layer(x, w_ih, w_hh, bias, lens_dev, tmax, save=True, x_bound=None):
    """x_bound: optional device scalar >= max|x| (the output of a previous layer is bounded by 1, or by 1/(1-p) after
    dropout), which saves the operand split its absmax pass."""
    return _BiLSTMLayer.apply(x, w_ih, w_hh, bias, lens_dev, tmax, save, x_bound)


# ---------------------------------------------------------------------------------------------------------------
# inter-layer LSTM dropout (reference nn.LSTM(dropout=p), cnnlstm.py:148-149)
# ---------------------------------------------------------------------------------------------------------------
_CONSTS = {}


def const_scalar(device, value):
    """A cached one-element fp32 device tensor (analytic operand bounds etc.)."""
    key = (str(device), float(value))
    t = _CONSTS.get(key)
    if t is None:
        t = _CONSTS[key] = torch.full((1,), float(value), dtype=F32, device=device)
    return t


def new_rng_state(device, seed=None, offset=0):
    """Device-resident Philox state {seed, offset} (int64[2]); the seed defaults to a draw from torch's CPU generator,
    so torch.manual_seed() makes dropout reproducible."""
    if seed is None:
        seed = int(torch.empty((), dtype=torch.int64).random_().item())
    return torch.tensor([int(seed) & 0x7FFFFFFFFFFFFFFF, int(offset)], dtype=torch.int64, device=device)


def rng_advance(rng, inc):
    check(lib().vocr_rng_advance(ptr(rng), int(inc), stream()), "vocr_rng_advance")


def dropout_keep_mask(n, p, seed, offset, device):
    """uint8[n] keep mask of the in-kernel Philox stream for (seed, offset) - what vocr_dropout_f32 applies."""
    m = torch.empty((n,), dtype=torch.uint8, device=device)
    st = lib().vocr_dropout_f32(None, None, n, float(p), None, int(seed), int(offset), None, None, ptr(m), stream())
    check(st, "vocr_dropout_f32")
    return m


class _Dropout(torch.autograd.Function):
    """y = x * keep / (1-p).  keep = `mask` (uint8, injected) or the Philox stream (rng state on the device + per-call
    `site` offset); backward re-applies the same mask without ever storing it."""

    @staticmethod
    def forward(ctx, x, p, rng, site, mask):
        x = _c(x)
        _lib.require_cuda(x, "x", F32)
        y = torch.empty_like(x)
        used = None
        if mask is not None:
            mask = _c(mask.to(device=x.device, dtype=torch.uint8))
            if mask.numel() != x.numel():
                raise _lib.VocrError("dropout mask has %d elements, input has %d" % (mask.numel(), x.numel()))
        else:
            used = torch.empty((2,), dtype=torch.int64, device=x.device)
        st = lib().vocr_dropout_f32(ptr(x), ptr(y), x.numel(), float(p), ptr(rng) if mask is None else None, 0,
                                    int(site), ptr(used), ptr(mask), None, stream())
        check(st, "vocr_dropout_f32")
        ctx.p = float(p)
        ctx.save_for_backward(used, mask)
        return y

    @staticmethod
    def backward(ctx, dy):
        used, mask = ctx.saved_tensors
        dy = _c(dy)
        dx = torch.empty_like(dy)
        st = lib().vocr_dropout_f32(ptr(dy), ptr(dx), dy.numel(), ctx.p, ptr(used), 0, 0, None, ptr(mask), None,
                                    stream())
        check(st, "vocr_dropout_f32")
        return dx, None, None, None, None


def dropout(x, p, rng=None, site=0, mask=None):
    if mask is None and rng is None:
        raise _lib.VocrError("dropout needs an rng state (ops.new_rng_state) or an injected mask")
    return _Dropout.apply(x, p, rng, site, mask)


# ---------------------------------------------------------------------------------------------------------------
# fused clamp + Adam over flat buffers
# ---------------------------------------------------------------------------------------------------------------
def clamp_adam_step(p, g, m, v, step, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, clamp=5.0,
                    grad_scale=1.0):
    for t, n in ((p, "p"), (g, "g"), (m, "m"), (v, "v")):
        _lib.require_cuda(t, n, F32)
    st = lib().vocr_clamp_adam_f32(ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), int(step), float(lr), float(betas[0]),
                                   float(betas[1]), float(eps), float(weight_decay), float(clamp), float(grad_scale),
                                   stream())
    check(st, "vocr_clamp_adam_f32")


def out_hw(h, w, n_rds):
    """Output (h, w) of rapid_ds + cnn for an input (h, w): reference cnn_input_size_to_output_size
    (cnnlstm.py:211-260): MaxPool2d floors the half, FractionalMaxPool2d floors x*0.5 / x*0.7 in float64."""
    for _ in range(n_rds):
        h, w = math.floor((h - 2) / 2 + 1), math.floor((w - 2) / 2 + 1)
    for _ in range(2):
        h, w = math.floor(h * 0.5), math.floor(w * 0.7)
    return h, w
