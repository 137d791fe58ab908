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


def _launch(x, b, xref, yref, dy, grad, dim, spec, alpha, gain, clamp):
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


def bias_act(x, b=None, dim=1, act="linear", alpha=None, gain=None, clamp=None, impl="cuda"):
    assert isinstance(x, torch.Tensor)
    if impl != "cuda":
        raise NotImplementedError("ic_gan_b200 has no PyTorch/CPU fallback for bias_act (impl='ref' lives in oracle/)")
    spec = activation_funcs[act]
    alpha = float(alpha if alpha is not None else spec.def_alpha)
    gain = float(gain if gain is not None else spec.def_gain)
    clamp = float(clamp if clamp is not None else -1)
    if b is not None:
        assert isinstance(b, torch.Tensor) and b.ndim == 1 and 0 <= dim < x.ndim and b.shape[0] == x.shape[dim]

    class BiasActCuda(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, b):
            y = x
            if act != "linear" or gain != 1 or clamp >= 0 or b is not None:
                y = _launch(x, b, None, None, None, 0, dim, spec, alpha, gain, clamp)
            ctx.save_for_backward(x if "x" in spec.ref or spec.has_2nd_grad else None,
                                  b if "x" in spec.ref or spec.has_2nd_grad else None,
                                  # the reference's CUDA path does not save y for act='linear' and therefore ignores the
                                  # clamp in backward (bias_act.py:236-241); its own impl='ref' (the pinned oracle) masks
                                  # it, which is the mathematically correct derivative -- followed here.
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
                db = dx.sum([i for i in range(dx.ndim) if i != dim])
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

    return BiasActCuda.apply(x, b)
