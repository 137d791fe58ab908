"""ctypes binding of libicgan_b200.so (the C ABI declared in include/icgan_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` (plain nvcc, sm_100a).  There is NO fallback: if the shared
object is missing or a call fails, a RuntimeError is raised (SURVEY.md §8b "Errors": no silent/CPU fallback).
PyTorch is used only for device memory and streams: every entry point receives raw device pointers and the current
CUDA stream handle.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libicgan_b200.so")

F32, BF16, F16 = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_TANH = 0, 1, 2

_lib: Optional[C.CDLL] = None

vp, fp, i32, i64, f32, f64 = C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double


class IcganSnLayer(C.Structure):
    _fields_ = [("W", C.c_void_p), ("u", C.c_void_p), ("v", C.c_void_p), ("u_new", C.c_void_p),
                ("sigma", C.c_void_p), ("scratch", C.c_void_p), ("rows", C.c_int), ("cols", C.c_int)]


# name -> argtypes (all return int). Mirrors include/icgan_b200.h one to one; tests check the two stay in sync.
SIGNATURES = {
    "icgan_conv2d_tc": [vp, vp, fp, fp, vp, vp, fp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp],
    "icgan_conv2d_tc_ex": [vp, vp, fp, fp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32,
                           i32, i32, i32, i32, i32, vp],
    "icgan_conv2d_wgrad_tc_ex": [vp, vp, fp, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp, i32, vp],
    "icgan_conv2d_rgb_tc": [vp, vp, fp, fp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp],
    "icgan_conv2d_wgrad_tc": [vp, vp, fp, i32, i32, i32, i32, i32, i32, vp],
    "icgan_conv2d_simt": [vp, fp, fp, fp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp],
    "icgan_conv2d_wgrad_simt": [vp, vp, fp, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp],
    "icgan_conv2d_small": [vp, fp, fp, fp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp],
    "icgan_conv2d_wgrad_small": [vp, vp, fp, i32, i32, i32, i32, i32, i32, i32, i32, vp],
    "icgan_im2col_small": [vp, vp, i32, i32, i32, i32, i32, i32, i32, vp],
    "icgan_channel_sum": [vp, fp, i64, i32, i32, vp],
    "icgan_nhwc_to_cnhw": [vp, vp, i64, i32, i32, vp],
    "icgan_sn_power_iteration": [vp, i32, i32, i32, i64, f32, i32, vp],
    "icgan_sn_prepare_weight": [fp, fp, vp, vp, i32, i32, i32, i32, vp],
    "icgan_sn_weight_grad": [fp, fp, fp, fp, fp, fp, fp, i32, i32, i32, vp],
    "icgan_bn_train_stats": [vp, i64, i32, i32, fp, fp, fp, fp, fp, fp, f32, f32, vp],
    "icgan_bn_stats_from_sums": [fp, fp, i64, i32, fp, fp, fp, fp, f32, f32, vp],
    "icgan_bn_apply": [vp, vp, fp, fp, fp, fp, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp],
    "icgan_bn_bwd_reduce": [vp, vp, fp, fp, fp, fp, i32, fp, fp, i32, i32, i32, i32, i32, i32, i32, i32, vp],
    "icgan_bn_bwd_apply": [vp, vp, vp, fp, fp, fp, fp, i32, fp, fp, i32, i32, i32, i32, i32, i32, i32, i32, vp],
    "icgan_relu": [vp, vp, i64, i32, vp],
    "icgan_relu_bwd": [vp, vp, vp, i64, i32, i32, vp],
    "icgan_tanh_bwd": [vp, vp, vp, i64, i32, i32, vp],
    "icgan_axpby": [vp, vp, vp, f32, fp, f32, fp, i64, i32, vp],
    "icgan_dot": [vp, vp, fp, i64, i32, vp],
    "icgan_pool2": [vp, vp, vp, i32, i32, i32, i32, f32, i32, i32, vp],
    "icgan_unpool2": [vp, vp, vp, i32, i32, i32, i32, f32, i32, i32, i32, vp],
    "icgan_relu_sumpool": [vp, fp, i32, i32, i32, i32, vp],
    "icgan_relu_sumpool_bwd": [vp, fp, vp, i32, i32, i32, i32, vp],
    "icgan_softmax_rows": [vp, vp, i64, i32, i32, i32, vp],
    "icgan_softmax_rows_bwd": [vp, vp, vp, i64, i32, i32, i32, i32, vp],
    "icgan_knn_prepare": [fp, vp, vp, fp, i64, i32, vp],
    "icgan_knn_coarse": [vp, vp, fp, i64, i32, i64, i64, i32, i32, vp, fp, vp],
    "icgan_knn_rerank": [fp, i64, i32, i64, i64, i32, i32, vp, fp, vp, vp, vp, fp, f32, vp],
    "icgan_knn_exact_row": [fp, i64, i32, i64, i32, vp, vp, vp, vp],
    "icgan_bias_act": [vp, vp, vp, vp, vp, vp, i64, i64, i32, i32, i32, f32, f32, f32, i32, vp],
    "icgan_upfirdn2d": [vp, fp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, f32, i32,
                        i32, vp],
    "icgan_upfirdn2d_nhwc": [vp, fp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, f32, fp, fp, fp, i32, fp, i32,
                             f32, f32, f32, fp, vp, vp, vp, i32, vp],
    "icgan_modulate": [vp, fp, vp, i32, i64, i32, i32, i32, vp],
    "icgan_chan_dot": [vp, vp, fp, i32, i64, i32, i32, i32, vp],
    "icgan_bias_act_nhwc": [vp, vp, vp, fp, fp, fp, fp, i32, i32, i64, i32, i32, i32, f32, f32, f32, i32, vp],
    "icgan_mod_bias_act_bwd": [vp, vp, vp, fp, vp, fp, fp, fp, i32, i64, i32, i32, f32, f32, f32, i32, vp],
    "icgan_adam_ema_step": [fp, fp, fp, fp, fp, i64, f64, f64, f64, f64, i64, f64, f64, vp],
    "icgan_ema_lerp": [fp, fp, i64, f64, vp],
    "icgan_gemm_tc": [vp, vp, vp, i32, i32, i32, i32, i32, i32, i64, i64, i64, i64, i64, i64, f32, i32, vp],
    "icgan_attn_fwd": [vp, vp, vp, vp, fp, i32, i32, i32, i32, i32, vp],
    "icgan_attn_bwd_q": [vp, vp, vp, vp, vp, fp, vp, vp, fp, i32, i32, i32, i32, i32, vp],
    "icgan_attn_bwd_kv": [vp, vp, vp, vp, fp, fp, vp, vp, i32, i32, i32, i32, i32, vp],
    "icgan_gemm": [vp, vp, vp, i32, i32, i32, i32, i64, i64, i64, i64, i64, i64, i64, i64, i64, f32, fp, f32, fp, i32,
                   i32, i32, vp],
}


def load() -> C.CDLL:
    """Load the shared library (once). Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build the CUDA library first (python -c 'import __graft_entry__ as g; g.build()'). "
            "ic_gan_b200 has no CPU or PyTorch fallback for its kernels.")
    lib = C.CDLL(LIB_PATH)
    lib.icgan_last_error.restype = C.c_char_p
    lib.icgan_version.restype = C.c_int
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.argtypes = argtypes
        fn.restype = C.c_int
    _lib = lib
    return lib


def last_error() -> str:
    return load().icgan_last_error().decode("utf-8", "replace")


# kernels launched per entry point (everything else launches exactly one); bench.py reports the running total
KERNELS_PER_CALL = {"icgan_mod_bias_act_bwd": 1, "icgan_bn_train_stats": 2, "icgan_sn_power_iteration": 5, "icgan_sn_weight_grad": 2, "icgan_knn_exact_row": 2}
LAUNCHES = 0


def call(name: str, *args) -> None:
    global LAUNCHES
    LAUNCHES += KERNELS_PER_CALL.get(name, 1)
    rc = getattr(load(), name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed (rc={rc}): {last_error()}")


def float_array(values):
    """Host float array argument."""
    return (C.c_float * len(values))(*[float(v) for v in values])


def int_array(values):
    """Host int array argument (the tap tables of icgan_conv2d_tc_ex / icgan_conv2d_wgrad_tc_ex)."""
    return (C.c_int * len(values))(*[int(v) for v in values])


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def dt(t) -> int:
    """dtype code of a tensor or torch.dtype."""
    d = t.dtype if isinstance(t, torch.Tensor) else t
    if d == torch.float32:
        return F32
    if d == torch.bfloat16:
        return BF16
    if d == torch.float16:
        return F16  # StyleGAN2 ops only
    raise TypeError(f"unsupported dtype {d}; activations must be float32 or bfloat16")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("ic_gan_b200 kernels need CUDA tensors (there is no CPU path)")
    if not (t.is_contiguous() or (t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last))):
        # the kernels index dense memory; a strided view (or a channels_last-converted OIHW weight) would be read as garbage
        raise RuntimeError(f"ic_gan_b200 kernels need dense tensors; got shape {tuple(t.shape)} strides {t.stride()}")
    return t.data_ptr()
