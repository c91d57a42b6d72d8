"""ctypes binding of the C ABI in include/vistaocr_b200.h.

There is NO fallback: if the shared library is missing this raises, and every op raises on a non-zero status.
Tensors cross the boundary as raw device pointers + the current CUDA stream.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libvistaocr_b200.so")

c_int = ctypes.c_int
c_f = ctypes.c_float
c_sz = ctypes.c_size_t
c_p = ctypes.c_void_p
c_ll = ctypes.c_longlong

# name -> (restype, argtypes); must list every symbol declared in include/vistaocr_b200.h
PROTOTYPES = {
    "vocr_version": (c_int, []),
    "vocr_status_string": (ctypes.c_char_p, [c_int]),
    "vocr_greedy_decode_f32": (c_int, [c_p, c_int, c_int, c_int, c_p, c_f, c_p, c_p, c_p, c_p, c_int, c_p]),
    "vocr_ctc_workspace_size": (c_sz, [c_int, c_int, c_int, c_int]),
    "vocr_ctc_loss_f32": (c_int, [c_p, c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_p, c_p, c_sz, c_p]),
    "vocr_gemm_f32": (c_int, [c_int, c_int, c_int, c_int, c_int, c_p, c_int, c_p, c_int, c_p, c_int, c_p, c_int,
                              c_int, c_p]),
    "vocr_colsum_f32": (c_int, [c_p, c_ll, c_int, c_int, c_p, c_int, c_p]),
    "vocr_conv_weight_layout_f32": (c_int, [c_p, c_int, c_int, c_p, c_p, c_p]),
    "vocr_conv3x3_fwd_f32": (c_int, [c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_p, c_p]),
    "vocr_conv3x3_wgrad_workspace_size": (c_sz, [c_int, c_int, c_int, c_int, c_int]),
    "vocr_conv3x3_wgrad_f32": (c_int, [c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_p, c_sz, c_p]),
    "vocr_rds_fwd_f32": (c_int, [c_p, c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_p]),
    "vocr_rds_unpool_f32": (c_int, [c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_p]),
    "vocr_bn_finalize_f32": (c_int, [c_p, c_ll, c_p, c_p, c_p, c_p, c_f, c_f, c_int, c_p, c_p, c_p, c_p, c_int, c_p]),
    "vocr_bn_relu_apply_f32": (c_int, [c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_ll, c_ll, c_ll, c_p]),
    "vocr_bn_relu_bwd_f32": (c_int, [c_p, c_p, c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_ll, c_ll,
                                     c_ll, c_p, c_p, c_p, c_p, c_p, c_p]),
    "vocr_fracpool_fwd_f32": (c_int, [c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_int, c_p]),
    "vocr_fracpool_bwd_f32": (c_int, [c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_int, c_p]),
    "vocr_bilstm_workspace_size": (c_sz, [c_int, c_int, c_int]),
    "vocr_bilstm_fwd_f32": (c_int, [c_p, c_p, c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_p, c_sz, c_p]),
    "vocr_bilstm_bwd_f32": (c_int, [c_p, c_p, c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_p, c_sz, c_p]),
    "vocr_clamp_adam_f32": (c_int, [c_p, c_p, c_p, c_p, c_ll, c_int, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_p]),
}

_lib = None


class VocrError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VocrError(
                "vistaocr_b200: %s is missing - build it with `python -m vistaocr_b200.build` "
                "(there is no CPU or PyTorch fallback for the hot path)" % LIB_PATH)
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(status, what):
    if status != 0:
        msg = lib().vocr_status_string(status).decode()
        raise VocrError("%s failed: %s (status %d)" % (what, msg, status))


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(t, name, dtype=None):
    if not t.is_cuda:
        raise VocrError("%s must be a CUDA tensor: the vistaocr_b200 hot path has no CPU fallback" % name)
    if dtype is not None and t.dtype != dtype:
        raise VocrError("%s must have dtype %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise VocrError("%s must be contiguous" % name)
    return t
