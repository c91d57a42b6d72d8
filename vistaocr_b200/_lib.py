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
c_ull = ctypes.c_ulonglong

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
    "vocr_conv3x3_fwd_f32": (c_int, [c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_p, c_p, c_p]),
    "vocr_conv3x3_wgrad_workspace_size": (c_sz, [c_int, c_int, c_int, c_int, c_int]),
    "vocr_conv3x3_wgrad_f32": (c_int, [c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_p, c_sz, c_p]),
    "vocr_rds_fwd_f32": (c_int, [c_p, c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_p]),
    "vocr_rds_unpool_f32": (c_int, [c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_p]),
    "vocr_rds_wgrad_c1_f32": (c_int, [c_p, c_p, c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_p, c_p]),
    "vocr_conv3x3_c16_fwd_f32": (c_int, [c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_p]),
    "vocr_conv3x3_c16_wgrad_f32": (c_int, [c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_p, c_p]),
    "vocr_bn_finalize_f32": (c_int, [c_p, c_ll, c_p, c_p, c_p, c_p, c_f, c_f, c_int, c_p, c_p, c_p, c_p, c_int, c_p,
                                     c_p, c_p]),
    "vocr_bn_relu_apply_f32": (c_int, [c_p, c_p, c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_ll, c_ll, c_ll,
                                       c_p, c_p, c_p, c_p, c_p]),
    "vocr_bn_relu_bwd_f32": (c_int, [c_p, c_p, c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_ll, c_ll,
                                     c_ll, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
    "vocr_fracpool_fwd_f32": (c_int, [c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_int, c_p]),
    "vocr_fracpool_bwd_f32": (c_int, [c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_int, c_p]),
    "vocr_bilstm_workspace_size": (c_sz, [c_int, c_int, c_int, c_int]),
    "vocr_bilstm_fwd_f32": (c_int, [c_p, c_p, c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_p, c_sz, c_p]),
    "vocr_bilstm_bwd_f32": (c_int, [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_p, c_sz, c_p]),
    "vocr_clamp_adam_f32": (c_int, [c_p, c_p, c_p, c_p, c_ll, c_int, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_p]),
    "vocr_dropout_f32": (c_int, [c_p, c_p, c_ll, c_f, c_p, c_ull, c_ull, c_p, c_p, c_p, c_p]),
    "vocr_rng_advance": (c_int, [c_p, c_ull, c_p]),
    "vocr_split_tf32_f32": (c_int, [c_p, c_p, c_p, c_ll, c_p]),
    "vocr_tc_conv3x3_fwd": (c_int, [c_p, c_p, c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_p]),
    "vocr_tc_conv3x3_wgrad_workspace_size": (c_sz, [c_int, c_int, c_int, c_int, c_int]),
    "vocr_tc_conv3x3_wgrad": (c_int, [c_p, c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_p, c_sz, c_p]),
    "vocr_colstats_f32": (c_int, [c_p, c_ll, c_int, c_p, c_p]),
    "vocr_lm_frontend_f32": (c_int, [c_p, c_int, c_int, c_int, c_p, c_p, c_p, c_int, ctypes.c_double, c_p, c_p]),
    "vocr_edit_distance_i32": (c_int, [c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_p, c_p]),
    "vocr_collate_lines_f32": (c_int, [c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_p, c_p, c_p, c_p, c_p, c_p]),
    "vocr_set_tc_products": (c_int, [c_int]),
    "vocr_get_tc_products": (c_int, []),
    "vocr_scale_lines_u8": (c_int, [c_p, c_p, c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_int, c_p, c_p]),
    "vocr_tc_gemm_tf32x3": (c_int, [c_int, c_int, c_int, c_int, c_int, c_p, c_p, c_int, c_p, c_p, c_int, c_p, c_int,
                                    c_p, c_int, c_int, c_p, c_sz, c_p]),
    "vocr_split_f16_f32": (c_int, [c_p, c_ll, c_p, c_p, c_p, c_p, c_p]),
    "vocr_im2col3x3_f16": (c_int, [c_p, c_int, c_int, c_int, c_int, c_p, c_p, c_p, c_p, c_p]),
    "vocr_tc_gemm_f16x3": (c_int, [c_int, c_int, c_int, c_int, c_int, c_p, c_p, c_int, c_p, c_p, c_p, c_int, c_p, c_p,
                                   c_int, c_p, c_int, c_int, c_p, c_sz, c_int, c_p]),
    "vocr_tc_conv3x3_fwd_f16": (c_int, [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_int,
                                        c_int, c_p, c_p]),
    "vocr_tc_conv3x3_wgrad_f16": (c_int, [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_p,
                                          c_sz, c_int, c_p]),
    "vocr_tc_gemm_f16x3_argmax": (c_int, [c_int, c_int, c_int, c_p, c_p, c_int, c_p, c_p, c_p, c_int, c_p, c_p, c_int, c_p,
                                          c_p, c_int, c_int, c_f, c_p, c_int, c_p]),
    "vocr_ctc_collapse_i32": (c_int, [c_p, c_int, c_int, c_p, c_p, c_p, c_p, c_int, c_p]),
    "vocr_bn_eval_bound_f32": (c_int, [c_p, c_p, c_p, c_p, c_p, c_int, c_int, c_p, c_p]),
    "vocr_tc_conv3x3_bnrelu_f16": (c_int, [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_ll, c_ll, c_ll, c_p, c_p,
                                           c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_int, c_p]),
}

_lib = None

# kernels launched per entry-point call (for bench.py's gpu_launches claim) and algorithmic work per call
# (flops for the GEMM-shaped kernels, bytes for the HBM-bound ones) as a function of the ctypes argument tuple
KERNELS_PER_CALL = {
    "vocr_greedy_decode_f32": 2, "vocr_ctc_loss_f32": 4, "vocr_gemm_f32": 1, "vocr_colsum_f32": 1,
    "vocr_conv_weight_layout_f32": 1, "vocr_conv3x3_fwd_f32": 1, "vocr_conv3x3_wgrad_f32": 2, "vocr_rds_fwd_f32": 1,
    "vocr_rds_unpool_f32": 1, "vocr_rds_wgrad_c1_f32": 2, "vocr_conv3x3_c16_fwd_f32": 1, "vocr_conv3x3_c16_wgrad_f32": 2, "vocr_bn_finalize_f32": 1, "vocr_bn_relu_apply_f32": 1, "vocr_bn_relu_bwd_f32": 5,
    "vocr_fracpool_fwd_f32": 1, "vocr_fracpool_bwd_f32": 4, "vocr_bilstm_fwd_f32": 1, "vocr_bilstm_bwd_f32": 1,
    "vocr_clamp_adam_f32": 1, "vocr_dropout_f32": 1, "vocr_rng_advance": 1, "vocr_split_tf32_f32": 1, "vocr_tc_gemm_tf32x3": 1,
    "vocr_split_f16_f32": 2, "vocr_im2col3x3_f16": 2, "vocr_tc_gemm_f16x3": 1, "vocr_tc_conv3x3_fwd_f16": 1, "vocr_tc_conv3x3_wgrad_f16": 2,
    "vocr_bn_eval_bound_f32": 1, "vocr_tc_conv3x3_bnrelu_f16": 1, "vocr_tc_gemm_f16x3_argmax": 1, "vocr_ctc_collapse_i32": 1,
    "vocr_tc_conv3x3_fwd": 1, "vocr_tc_conv3x3_wgrad": 2, "vocr_colstats_f32": 1, "vocr_collate_lines_f32": 2, "vocr_scale_lines_u8": 1, "vocr_lm_frontend_f32": 1, "vocr_edit_distance_i32": 1,
}
WORK = {
    "vocr_gemm_f32": lambda a: ("flop", 2.0 * a[2] * a[3] * a[4]),
    "vocr_tc_gemm_tf32x3": lambda a: ("flop", 2.0 * a[2] * a[3] * a[4]),
    "vocr_split_tf32_f32": lambda a: ("byte", 12.0 * a[3]),
    "vocr_tc_gemm_f16x3": lambda a: ("flop", 2.0 * a[2] * a[3] * a[4]),
    "vocr_split_f16_f32": lambda a: ("byte", (8.0 if a[2] else 12.0) * a[1]),
    "vocr_im2col3x3_f16": lambda a: ("byte", 4.0 * a[1] * a[2] * a[3] * a[4] * 10.0),
    "vocr_tc_conv3x3_fwd_f16": lambda a: ("flop", 2.0 * a[8] * a[9] * a[10] * 9 * a[11] * a[12]),
    "vocr_tc_conv3x3_wgrad_f16": lambda a: ("flop", 2.0 * a[7] * a[8] * a[9] * 9 * a[10] * a[11]),
    "vocr_tc_gemm_f16x3_argmax": lambda a: ("flop", 2.0 * a[0] * a[1] * a[2]),
    "vocr_tc_conv3x3_bnrelu_f16": lambda a: ("flop", 2.0 * a[18] * a[19] * a[20] * 9 * a[21] * a[22]),
    "vocr_tc_conv3x3_fwd": lambda a: ("flop", 2.0 * a[6] * a[7] * a[8] * 9 * a[9] * a[10]),
    "vocr_tc_conv3x3_wgrad": lambda a: ("flop", 2.0 * a[5] * a[6] * a[7] * 9 * a[8] * a[9]),
    "vocr_conv3x3_fwd_f32": lambda a: ("flop", 2.0 * a[4] * a[5] * a[6] * 9 * a[7] * a[8]),
    "vocr_conv3x3_c16_fwd_f32": lambda a: ("flop", 2.0 * a[4] * a[5] * a[6] * 9 * 16 * 16),
    "vocr_conv3x3_c16_wgrad_f32": lambda a: ("flop", 2.0 * a[4] * a[5] * a[6] * 9 * 16 * 16),
    "vocr_conv3x3_wgrad_f32": lambda a: ("flop", 2.0 * a[3] * a[4] * a[5] * 9 * a[6] * a[7]),
    # the recurrent step is HBM / latency bound; per (direction, step) it reads the x-projection
    # (B*4H) and writes h (B*H) - backward: reads dout, the gates, c and writes the gate gradients (B*10H) - W_hh and the
    # running state stay on chip
    "vocr_bilstm_fwd_f32": lambda a: ("byte", 4.0 * a[9] * 2 * a[7] * 5 * a[8]),
    "vocr_bilstm_bwd_f32": lambda a: ("byte", 4.0 * a[10] * 2 * a[8] * 10 * a[9]),
    "vocr_greedy_decode_f32": lambda a: ("byte", 4.0 * a[1] * a[2] * a[3]),
    "vocr_ctc_loss_f32": lambda a: ("byte", 8.0 * a[5] * a[6] * a[7]),
    "vocr_clamp_adam_f32": lambda a: ("byte", 28.0 * a[4]),
    "vocr_dropout_f32": lambda a: ("byte", 8.0 * a[2]),
}


class Profiler:
    """Counts entry-point calls / kernel launches and, when `timing` is on, brackets every call with CUDA events on
    the launching stream (bench.py reads per-entry-point device time and work from here)."""

    def __init__(self):
        self.reset()
        self.timing = False

    def reset(self):
        self.calls = {}
        self.launches = 0
        self.records = []  # (name, start_event, end_event, kind, work)
        self.arg_log = []  # integer arguments of every recorded call (when keep_args is set; tools/step_profile.py)
        self.keep_args = getattr(self, "keep_args", False)

    def summary(self):
        out = {}
        for name, s, e, kind, work in self.records:
            d = out.setdefault(name, {"ms": 0.0, "calls": 0, "kind": kind, "work": 0.0})
            d["ms"] += s.elapsed_time(e)
            d["calls"] += 1
            d["work"] += work
        return out


PROFILER = Profiler()


def _wrap(name, fn):
    k = KERNELS_PER_CALL.get(name)
    if k is None:
        return fn
    work_fn = WORK.get(name)

    def call(*args):
        p = PROFILER
        p.calls[name] = p.calls.get(name, 0) + 1
        p.launches += k
        if not p.timing:
            return fn(*args)
        s = torch.cuda.Event(enable_timing=True)
        e = torch.cuda.Event(enable_timing=True)
        s.record()
        st = fn(*args)
        e.record()
        kind, work = work_fn(args) if work_fn else ("none", 0.0)
        p.records.append((name, s, e, kind, work))
        if p.keep_args:
            p.arg_log.append(tuple(a for a in args if isinstance(a, int)))
        return st

    return call


class _Lib:
    pass


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
        w = _Lib()
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
            setattr(w, name, _wrap(name, fn))
        _lib = w
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
