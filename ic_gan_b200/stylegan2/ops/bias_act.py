"""Drop-in for ``stylegan2_ada_pytorch/torch_utils/ops/bias_act.py`` (public function ``bias_act`` :131-171).

Same signature, activation table, defaults and derivative structure (first order for every activation, second order for
the ones with ``has_2nd_grad``), running on ``icgan_bias_act``.  ``impl='ref'`` of the reference is a PyTorch fallback;
this package has none (the CPU restatement lives in oracle/ and is test-only)."""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np
import torch

from ..._lib import call, dt, ptr, stream_ptr

# name -> (def_alpha, def_gain, cuda_idx, ref, has_2nd_grad)   (bias_act.py:26-99)
activation_funcs = {
    "linear": SimpleNamespace(def_alpha=0, def_gain=1, cuda_idx=1, ref="", has_2nd_grad=False),
    "relu": SimpleNamespace(def_alpha=0, def_gain=math.sqrt(2), cuda_idx=2, ref="y", has_2nd_grad=False),
    "lrelu": SimpleNamespace(def_alpha=0.2, def_gain=math.sqrt(2), cuda_idx=3, ref="y", has_2nd_grad=False),
    "tanh": SimpleNamespace(def_alpha=0, def_gain=1, cuda_idx=4, ref="y", has_2nd_grad=True),
    "sigmoid": SimpleNamespace(def_alpha=0, def_gain=1, cuda_idx=5, ref="y", has_2nd_grad=True),
    "elu": SimpleNamespace(def_alpha=0, def_gain=1, cuda_idx=6, ref="y", has_2nd_grad=True),
    "selu": SimpleNamespace(def_alpha=0, def_gain=1, cuda_idx=7, ref="y", has_2nd_grad=True),
    "softplus": SimpleNamespace(def_alpha=0, def_gain=1, cuda_idx=8, ref="y", has_2nd_grad=True),
    "swish": SimpleNamespace(def_alpha=0, def_gain=math.sqrt(2), cuda_idx=9, ref="x", has_2nd_grad=True),
}


def _fast_nhwc(x, b, yref, grad, dim, spec, alpha, gain, clamp):
    """Channels-last 4-D tensors with lrelu / linear (everything the training path issues): the vectorised kernel
    (8 elements per thread, no per-element div/mod).  Returns None when the call is not eligible."""
    if not (x.ndim == 4 and dim == 1 and spec.cuda_idx in (1, 3) and grad in (0, 1) and x.shape[1] % 8 == 0
            and x.shape[1] > 1 and x.dtype in (torch.float32, torch.bfloat16)
            and x.is_contiguous(memory_format=torch.channels_last)):
        return None
    if grad == 1 and yref is None and not (spec.cuda_idx == 1 and clamp < 0):
        return None
    N, C, H, W = x.shape
    y = torch.empty_like(x)  # preserves channels-last
    bb = None if (b is None or grad == 1) else b.float().contiguous()
    yr = None if yref is None else yref.to(x.dtype).contiguous(memory_format=torch.channels_last)
    call("icgan_bias_act_nhwc", ptr(x), ptr(yr), ptr(y), ptr(bb), None, None, None, 0, N, H * W, C, grad, spec.cuda_idx,
         float(alpha), float(gain), float(clamp), dt(x), stream_ptr())
    return y


def _launch(x, b, xref, yref, dy, grad, dim, spec, alpha, gain, clamp):
    if dy is None:
        fast = _fast_nhwc(x, b, yref, grad, dim, spec, alpha, gain, clamp)
        if fast is not None:
            return fast
    x = x.contiguous(memory_format=torch.channels_last) if (x.ndim == 4 and x.stride(1) == 1 and x.shape[1] > 1) \
        else x.contiguous()
    y = torch.empty_like(x)
    # bias index of flat element i: (i / step_b) % size_b with step_b from the actual strides (dense NCHW or NHWC)
    step_b = x.stride(dim) if x.ndim > 0 else 1
    size_b = x.shape[dim] if b is not None else 1

    def same(t):
        if t is None:
            return None
        t = t.to(x.dtype)
        return t.contiguous(memory_format=torch.channels_last) if (x.ndim == 4 and x.stride(1) == 1 and x.shape[1] > 1) \
            else t.contiguous()
    bb = None if b is None else b.to(x.dtype).contiguous()
    xr, yr, dd = same(xref), same(yref), same(dy)  # keep the (possibly re-laid-out) copies alive across the launch
    call("icgan_bias_act", ptr(x), ptr(bb), ptr(xr), ptr(yr), ptr(dd), ptr(y), x.numel(), int(step_b), int(size_b), grad,
         spec.cuda_idx, float(alpha), float(gain), float(clamp), dt(x), stream_ptr())
    del xr, yr, dd
    return y


_fn_cache = {}


def _function(dim, act, alpha, gain, clamp):
    """The autograd Function of one (dim, act, alpha, gain, clamp) configuration, built once (a training step issues the
    same handful of configurations thousands of times)."""
    key = (dim, act, alpha, gain, clamp)
    if key in _fn_cache:
        return _fn_cache[key]
    spec = activation_funcs[act]
    needs_x = "x" in spec.ref or spec.has_2nd_grad

    class BiasActCuda(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, b):
            y = x
            if act != "linear" or gain != 1 or clamp >= 0 or b is not None:
                y = _launch(x, b, None, None, None, 0, dim, spec, alpha, gain, clamp)
            # The reference's CUDA path does not save y for act='linear' and therefore ignores the clamp in backward
            # (bias_act.py:236-241); its own impl='ref' (the pinned oracle) masks it, which is the mathematically
            # correct derivative -- followed here.
            ctx.save_for_backward(x if needs_x else None, b if needs_x else None,
                                  y if ("y" in spec.ref or clamp >= 0) else None)
            ctx.has_b = b is not None
            return y

        @staticmethod
        def backward(ctx, dy):
            x, b, y = ctx.saved_tensors
            dx = db = None
            if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
                dx = dy
                if act != "linear" or gain != 1 or clamp >= 0:
                    dx = BiasActCudaGrad.apply(dy, x, b, y)
            if ctx.needs_input_grad[1] and ctx.has_b:
                db = dx.sum([i for i in range(dx.ndim) if i != dim], dtype=torch.float32).to(dx.dtype)
            return dx, db

    class BiasActCudaGrad(torch.autograd.Function):
        @staticmethod
        def forward(ctx, dy, x, b, y):
            dx = _launch(dy, b, x, y, None, 1, dim, spec, alpha, gain, clamp)
            ctx.save_for_backward(dy if spec.has_2nd_grad else None, x, b, y)
            return dx

        @staticmethod
        def backward(ctx, d_dx):
            dy, x, b, y = ctx.saved_tensors
            d_dy = d_x = d_b = None
            if ctx.needs_input_grad[0]:
                d_dy = BiasActCudaGrad.apply(d_dx, x, b, y)
            if spec.has_2nd_grad and (ctx.needs_input_grad[1] or ctx.needs_input_grad[2]):
                d_x = _launch(d_dx, b, x, y, dy, 2, dim, spec, alpha, gain, clamp)
            if spec.has_2nd_grad and ctx.needs_input_grad[2]:
                d_b = d_x.sum([i for i in range(d_x.ndim) if i != dim])
            return d_dy, d_x, d_b, None

    _fn_cache[key] = BiasActCuda
    return BiasActCuda


def bias_act(x, b=None, dim=1, act="linear", alpha=None, gain=None, clamp=None, impl="cuda"):
    if not isinstance(x, torch.Tensor):
        raise TypeError("bias_act: x must be a tensor")
    if impl != "cuda":
        raise NotImplementedError("ic_gan_b200 has no PyTorch/CPU fallback for bias_act (impl='ref' lives in oracle/)")
    spec = activation_funcs[act]
    alpha = float(alpha if alpha is not None else spec.def_alpha)
    gain = float(gain if gain is not None else spec.def_gain)
    clamp = float(clamp if clamp is not None else -1)
    if b is not None and not (isinstance(b, torch.Tensor) and b.ndim == 1 and 0 <= dim < x.ndim and b.shape[0] == x.shape[dim]):
        raise ValueError("bias_act: b must be a vector matching x.shape[dim]")
    return _function(dim, act, alpha, gain, clamp).apply(x, b)
