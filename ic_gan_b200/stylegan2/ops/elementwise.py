"""The elementwise pieces of StyleGAN2's modulated convolution as channels-last kernels with closed-form derivatives.

``modulated_conv2d`` in training mode is ``conv(x * styles[n,ci]) * dcoefs[n,co] + noise`` followed by ``bias_act``
(stylegan2_ada_pytorch/training/networks.py:77-95, :441-444).  Around the convolution that is five elementwise passes in
the reference (mul, fma, bias_act and their backward counterparts); here it is

* ``modulate(x, s)``            -- ``x * s[n,c]``, one pass, optional float32 -> bfloat16 cast on the way out;
* ``chan_dot(a, b)``            -- ``sum_hw a*b -> [n,c]``, the adjoint of ``modulate`` (gradients w.r.t. styles / dcoefs);
* ``mod_bias_act(x, pre, noise, bias, act, gain, clamp)`` -- ``clamp(act(x*pre[n,c] + noise[n,hw] + bias[c]) * gain)``,
  one pass for demodulation + noise + bias + activation + gain + clamp.

Every backward below is written with these same Functions (and reductions), so gradients of gradients -- the path-length
regulariser differentiates the synthesis network twice -- need no extra code."""
from __future__ import annotations

import torch

from ..._lib import call, dt, ptr, stream_ptr

_ACT_ID = {"linear": 1, "lrelu": 3}


def _cl(x):
    return x.contiguous(memory_format=torch.channels_last)


class ModulateFn(torch.autograd.Function):
    """y[n,c,h,w] = x[n,c,h,w] * s[n,c]; x channels-last float32 / bfloat16, s float32 [N,C]."""

    @staticmethod
    def forward(ctx, x, s, out_dtype):
        # the INPUT tensors are saved (not the re-laid-out copies): a second differentiation must reach their history
        ctx.save_for_backward(x, s)
        ctx.in_dtype = x.dtype
        xc, sc = _cl(x), s.float().contiguous()
        N, C, H, W = x.shape
        y = torch.empty((N, C, H, W), device=x.device, dtype=out_dtype or x.dtype, memory_format=torch.channels_last)
        call("icgan_modulate", ptr(xc), ptr(sc), ptr(y), N, H * W, C, dt(xc), dt(y), stream_ptr())
        return y

    @staticmethod
    def backward(ctx, dy):
        x, s = ctx.saved_tensors
        dx = ds = None
        if ctx.needs_input_grad[0]:
            dx = ModulateFn.apply(dy, s, ctx.in_dtype)
        if ctx.needs_input_grad[1]:
            ds = ChanDotFn.apply(dy, x)
        return dx, ds, None


class ChanDotFn(torch.autograd.Function):
    """out[n,c] = sum_hw a[n,c,h,w] * b[n,c,h,w]  (float32)."""

    @staticmethod
    def forward(ctx, a, b):
        ctx.save_for_backward(a, b)
        ac, bc = _cl(a), _cl(b)
        N, C, H, W = a.shape
        out = torch.empty(N, C, device=a.device, dtype=torch.float32)
        call("icgan_chan_dot", ptr(ac), ptr(bc), ptr(out), N, H * W, C, dt(ac), dt(bc), stream_ptr())
        return out

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        da = ModulateFn.apply(b, g, a.dtype) if ctx.needs_input_grad[0] else None
        db = ModulateFn.apply(a, g, b.dtype) if ctx.needs_input_grad[1] else None
        return da, db


def modulate(x, s, out_dtype=None):
    if x.shape[1] % 8 == 0 and x.dtype in (torch.float32, torch.bfloat16):
        return ModulateFn.apply(x, s, out_dtype)
    y = x * s.to(x.dtype).reshape(x.shape[0], -1, 1, 1)  # RGB-sized channel counts
    return y if out_dtype is None else y.to(out_dtype)


def chan_dot(a, b):
    if a.shape[1] % 8 == 0 and a.dtype in (torch.float32, torch.bfloat16) and b.dtype in (torch.float32, torch.bfloat16):
        return ChanDotFn.apply(a, b)
    return (a.float() * b.float()).sum([2, 3])


class _ActGradFn(torch.autograd.Function):
    """t = dy * gain * act'(y) * [|y| < clamp]  for act in {linear, lrelu}: linear in dy, piecewise constant in y."""

    @staticmethod
    def forward(ctx, dy, y, act_id, alpha, gain, clamp):
        dy = _cl(dy)
        if y is not None and dy.dtype != y.dtype:
            dy = dy.to(y.dtype)
        N, C, H, W = dy.shape
        t = torch.empty_like(dy)
        call("icgan_bias_act_nhwc", ptr(dy), ptr(y), ptr(t), None, None, None, None, 0, N, H * W, C, 1, act_id, float(alpha),
             float(gain), float(clamp), dt(dy), stream_ptr())
        ctx.save_for_backward(y)
        ctx.cfg = (act_id, alpha, gain, clamp)
        return t

    @staticmethod
    def backward(ctx, dt_):
        (y,) = ctx.saved_tensors
        return _ActGradFn.apply(dt_, y, *ctx.cfg), None, None, None, None, None


class ModBiasActFn(torch.autograd.Function):
    """y = clamp(act(x * pre[n,c] + noise[n or 1, 1, h, w] + bias[c]) * gain); x channels-last, C % 8 == 0."""

    @staticmethod
    def forward(ctx, x, pre, noise, bias, act_id, alpha, gain, clamp):
        xc = _cl(x)
        N, C, H, W = x.shape
        pre_c = None if pre is None else pre.float().contiguous()
        nz = None if noise is None else noise.float().reshape(-1, H, W).contiguous()
        b = None if bias is None else bias.float().contiguous()
        y = torch.empty_like(xc)
        call("icgan_bias_act_nhwc", ptr(xc), None, ptr(y), ptr(b), ptr(pre_c), ptr(nz), None,
             int(nz is not None and nz.shape[0] > 1), N, H * W, C, 0, act_id, float(alpha), float(gain), float(clamp),
             dt(xc), stream_ptr())
        # inputs (and the output) themselves, so second derivatives reach their history; a linear, unclamped epilogue needs
        # no y in its backward, which keeps the caller free to update y in place (`y.add_(x)` of the residual blocks)
        needs_y = act_id != 1 or clamp >= 0
        ctx.save_for_backward(x, pre, y if needs_y else None)
        ctx.shape, ctx.out_dtype = tuple(y.shape), y.dtype
        ctx.cfg = (act_id, alpha, gain, clamp)
        ctx.noise_shape = None if noise is None else tuple(noise.shape)
        ctx.noise_dtype = None if noise is None else noise.dtype
        ctx.bias_dtype = None if bias is None else bias.dtype
        return y

    @staticmethod
    def backward(ctx, dy):
        x, pre, y = ctx.saved_tensors
        if pre is not None:
            pre = pre.float()
        N, C, H, W = ctx.shape
        if (not torch.is_grad_enabled() and ctx.out_dtype == torch.bfloat16 and x.dtype == torch.bfloat16
                and 256 % (C // 8) == 0 and x.is_contiguous(memory_format=torch.channels_last)):
            # first-order backward (no create_graph): everything in ONE pass over dy / y / x
            act_id, alpha, gain, clamp = ctx.cfg
            dyc = _cl(dy).to(torch.bfloat16)
            need_pre = pre is not None and ctx.needs_input_grad[1]
            dev = dyc.device
            dx = torch.empty_like(dyc)
            dpre = torch.empty(N, C, device=dev, dtype=torch.float32) if need_pre else None
            db_n = torch.empty(N, C, device=dev, dtype=torch.float32) if (ctx.bias_dtype is not None and ctx.needs_input_grad[3]) else None
            dnz = torch.empty(N, H, W, device=dev, dtype=torch.float32) if (ctx.noise_shape is not None and ctx.needs_input_grad[2]) else None
            pre_c = None if pre is None else pre.contiguous()
            call("icgan_mod_bias_act_bwd", ptr(dyc), ptr(y if y is not None else dyc), ptr(x) if need_pre else None, ptr(pre_c),
                 ptr(dx), ptr(dpre), ptr(db_n), ptr(dnz), N, H * W, C, act_id, float(alpha), float(gain), float(clamp),
                 dt(dyc), stream_ptr())
            dnoise = dbias = None
            if dnz is not None:
                dnoise = dnz.sum(0, keepdim=True) if ctx.noise_shape[0] == 1 and N != 1 else dnz
                dnoise = dnoise.reshape(ctx.noise_shape).to(ctx.noise_dtype)
            if db_n is not None:
                dbias = db_n.sum(0).to(ctx.bias_dtype)
            return (dx if ctx.needs_input_grad[0] else None), dpre, dnoise, dbias, None, None, None, None
        t = _ActGradFn.apply(dy, y, *ctx.cfg)  # gradient w.r.t. the pre-activation
        dx = dpre = dnoise = dbias = None
        if ctx.needs_input_grad[0]:
            dx = t if pre is None else ModulateFn.apply(t, pre, x.dtype)
        if pre is not None and ctx.needs_input_grad[1]:
            dpre = ChanDotFn.apply(t, x)
        if ctx.noise_shape is not None and ctx.needs_input_grad[2]:
            dnoise = t.sum(dim=1, keepdim=True, dtype=torch.float32)  # fp32 accumulate, no widened copy of t
            if ctx.noise_shape[0] == 1 and dnoise.shape[0] != 1:
                dnoise = dnoise.sum(dim=0, keepdim=True)
            dnoise = dnoise.reshape(ctx.noise_shape).to(ctx.noise_dtype)
        if ctx.bias_dtype is not None and ctx.needs_input_grad[3]:
            dbias = t.sum([0, 2, 3], dtype=torch.float32).to(ctx.bias_dtype)
        return dx, dpre, dnoise, dbias, None, None, None, None


def mod_bias_act(x, pre=None, noise=None, bias=None, act="linear", alpha=0.2, gain=1.0, clamp=None):
    """See the module docstring; `noise` is [N or 1, 1, H, W] (already multiplied by its strength) or None."""
    c = -1.0 if clamp is None else float(clamp)
    if x.ndim == 4 and x.shape[1] % 8 == 0 and x.dtype in (torch.float32, torch.bfloat16) and act in _ACT_ID:
        if noise is not None and noise.ndim != 4:
            noise = noise.reshape(-1, 1, x.shape[2], x.shape[3])
        return ModBiasActFn.apply(x, pre, noise, bias, _ACT_ID[act], alpha, float(gain), c)
    # RGB-sized channel counts / other activations: the same function from the general ops
    from . import bias_act as ba
    if pre is not None:
        x = x * pre.to(x.dtype).reshape(x.shape[0], -1, 1, 1)
    if noise is not None:
        x = x + noise.to(x.dtype).reshape(-1, 1, x.shape[2], x.shape[3])
    return ba.bias_act(x, None if bias is None else bias.to(x.dtype), act=act, gain=gain, clamp=clamp)
