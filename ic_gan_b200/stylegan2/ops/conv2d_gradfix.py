"""Drop-in for ``stylegan2_ada_pytorch/torch_utils/ops/conv2d_gradfix.py``: ``conv2d`` (:43-62), ``conv_transpose2d``
(:65-94), ``no_weight_gradients`` (:31-37), differentiable to arbitrary order (R1 and path-length regularisation need
grad-of-grad through the convolutions; reference :126-272).  Runs on libicgan_b200: tcgen05 implicit GEMM where the shape
allows (bf16, 1x1/3x3, stride 1, 'same' padding), the CUDA-core NHWC kernels otherwise (any stride/padding, fp32
accumulate); transposed convolution = zero-insertion (icgan_upfirdn2d) + stride-1 convolution with the flipped kernel."""
from __future__ import annotations

import contextlib

import torch

from ... import _lib as L
from ..._lib import call, dt, ptr, stream_ptr

weight_gradients_disabled = False


@contextlib.contextmanager
def no_weight_gradients():
    global weight_gradients_disabled
    old = weight_gradients_disabled
    weight_gradients_disabled = True
    try:
        yield
    finally:
        weight_gradients_disabled = old


def _pair(v):
    return (int(v), int(v)) if isinstance(v, int) else (int(v[0]), int(v[1]))


def _square(name, v):
    a, b = _pair(v)
    if a != b:
        raise NotImplementedError(f"ic_gan_b200 conv2d_gradfix: anisotropic {name}={v} is not implemented")
    return a


def _nhwc(x):
    work = torch.float32 if x.dtype == torch.float16 else x.dtype
    return x.permute(0, 2, 3, 1).contiguous().to(work)


def _conv_nhwc(x, wk32, bias, stride, pad):
    """x [B,H,W,Ci] (f32/bf16), wk32 [Co,k,k,Ci] float32 -> [B,Ho,Wo,Co] in x.dtype."""
    B, H, W, ci = x.shape
    co, k = wk32.shape[0], wk32.shape[1]
    ho, wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    y = torch.empty(B, ho, wo, co, device=x.device, dtype=x.dtype)
    b32 = None if bias is None else bias.float().contiguous()
    if x.dtype == torch.bfloat16 and stride == 1 and k in (1, 3) and pad == k // 2 and ci % 16 == 0 and co % 8 == 0:
        wk16 = wk32.to(torch.bfloat16)
        call("icgan_conv2d_tc", ptr(x), ptr(wk16), None, ptr(b32), None, ptr(y), None, B, H, W, ci, co,
             k, dt(y), L.F32, 0, L.ACT_NONE, stream_ptr())
    else:
        call("icgan_conv2d_simt", ptr(x), ptr(wk32), None, ptr(b32), None, ptr(y), B, H, W, ci, co, k, stride, pad, dt(x),
             dt(y), L.F32, 0, L.ACT_NONE, stream_ptr())
    return y


def _zero_insert(x, stride, pad0, pad1):
    """NHWC zero-insertion upsampling by `stride` with (possibly negative) padding, via the upfirdn2d kernel."""
    B, H, W, Cc = x.shape
    one = torch.ones(1, 1, device=x.device, dtype=torch.float32)
    oh, ow = H * stride + pad0 + pad1, W * stride + pad0 + pad1
    y = torch.empty(B, oh, ow, Cc, device=x.device, dtype=x.dtype)
    call("icgan_upfirdn2d", ptr(x), ptr(one), ptr(y), B, Cc, H, W, 1, 1, stride, stride, 1, 1, pad0, pad1, pad0, pad1, 0,
         1.0, 1, dt(x), stream_ptr())
    return y


def _conv_core(x, w, b, stride, pad):
    """conv2d, NCHW in/out: y = corr(x, w) + b."""
    xin = _nhwc(x)
    wk = w.float().permute(0, 2, 3, 1).contiguous()
    y = _conv_nhwc(xin, wk, b, stride, pad)
    return y.permute(0, 3, 1, 2).to(x.dtype)


def _conv_transpose_core(x, w, b, stride, pad, out_pad):
    """conv_transpose2d, NCHW in/out; w is [Cin, Cout, k, k]."""
    xin = _nhwc(x)
    k = w.shape[2]
    q = k - 1 - pad
    if stride > 1 or q != 0 or out_pad != 0:
        xin = _zero_insert(xin, stride, q, q + out_pad - (stride - 1))
    wk = w.float().flip([2, 3]).permute(1, 2, 3, 0).contiguous()  # [Cout, k, k, Cin], flipped taps
    y = _conv_nhwc(xin, wk, b, 1, 0)
    return y.permute(0, 3, 1, 2).to(x.dtype)


def _wgrad_core(x, dy, stride, pad, k):
    """dW[o,i,kh,kw] = sum dy[n,o,ho,wo] x[n,i,ho*s+kh-p,wo*s+kw-p]  (float32, OIHW)."""
    xin, gin = _nhwc(x), _nhwc(dy)
    if gin.dtype != xin.dtype:
        gin = gin.to(xin.dtype)
    B, H, W, ci = xin.shape
    co = gin.shape[3]
    G = torch.zeros(co, k, k, ci, device=x.device, dtype=torch.float32)
    if xin.dtype == torch.bfloat16 and stride == 1 and k in (1, 3) and pad == k // 2 and ci % 16 == 0 and co % 8 == 0:
        call("icgan_conv2d_wgrad_tc", ptr(xin), ptr(gin), ptr(G), B, H, W, ci, co, k, stream_ptr())
    else:
        call("icgan_conv2d_wgrad_simt", ptr(xin), ptr(gin), ptr(G), B, H, W, ci, co, k, stride, pad, dt(xin), stream_ptr())
    return G.permute(0, 3, 1, 2).contiguous()


_cache = {}


def _op(transpose, weight_shape, stride, padding, output_padding, groups):
    key = (transpose, tuple(weight_shape), stride, padding, output_padding, groups)
    if key in _cache:
        return _cache[key]
    k = weight_shape[2]
    if weight_shape[2] != weight_shape[3]:
        raise NotImplementedError("ic_gan_b200 conv2d_gradfix: square kernels only")

    def out_pad_for(input_shape, output_shape):
        if transpose:
            return 0
        return input_shape[2] - (output_shape[2] - 1) * stride - (1 - 2 * padding) - (k - 1)

    def run(x, w, b):
        if groups == 1:
            return (_conv_transpose_core(x, w, b, stride, padding, output_padding) if transpose
                    else _conv_core(x, w, b, stride, padding))
        xs = x.chunk(groups, 1)
        ws = w.chunk(groups, 0)
        bs = [None] * groups if b is None else b.chunk(groups, 0)
        return torch.cat([(_conv_transpose_core(xi, wi, bi, stride, padding, output_padding) if transpose
                           else _conv_core(xi, wi, bi, stride, padding)) for xi, wi, bi in zip(xs, ws, bs)], 1)

    class Conv2d(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, w, b):
            ctx.save_for_backward(x, w)
            ctx.has_b = b is not None
            return run(x, w, b)

        @staticmethod
        def backward(ctx, dy):
            x, w = ctx.saved_tensors
            dx = dw = db = None
            if ctx.needs_input_grad[0]:
                p = out_pad_for(x.shape, dy.shape)
                dx = _op(not transpose, weight_shape, stride, padding, p, groups).apply(dy, w, None)
            if ctx.needs_input_grad[1] and not weight_gradients_disabled:
                dw = Conv2dGradWeight.apply(dy, x)
            if ctx.needs_input_grad[2] and ctx.has_b:
                db = dy.float().sum([0, 2, 3]).to(dy.dtype)
            return dx, dw, db

    class Conv2dGradWeight(torch.autograd.Function):
        @staticmethod
        def forward(ctx, dy, x):
            ctx.save_for_backward(dy, x)
            a, g = (x, dy) if not transpose else (dy, x)  # transposed conv: the roles of input and output swap
            if groups == 1:
                dw = _wgrad_core(a, g, stride, padding, k)
            else:
                dw = torch.cat([_wgrad_core(ai, gi, stride, padding, k)
                                for ai, gi in zip(a.chunk(groups, 1), g.chunk(groups, 1))], 0)
            return dw.to(x.dtype) if x.dtype != torch.float16 else dw.to(torch.float16)

        @staticmethod
        def backward(ctx, d2w):
            dy, x = ctx.saved_tensors
            d2y = d2x = None
            if ctx.needs_input_grad[0]:
                d2y = Conv2d.apply(x, d2w, None)
            if ctx.needs_input_grad[1]:
                p = out_pad_for(x.shape, dy.shape)
                d2x = _op(not transpose, weight_shape, stride, padding, p, groups).apply(dy, d2w, None)
            return d2y, d2x

    _cache[key] = Conv2d
    return Conv2d


def conv2d(input, weight, bias=None, stride=1, padding=0, dilation=1, groups=1):
    if _pair(dilation) != (1, 1):
        raise NotImplementedError("ic_gan_b200 conv2d_gradfix: dilation is not implemented (unused by StyleGAN2)")
    return _op(False, weight.shape, _square("stride", stride), _square("padding", padding), 0, groups).apply(
        input, weight, bias)


def conv_transpose2d(input, weight, bias=None, stride=1, padding=0, output_padding=0, groups=1, dilation=1):
    if _pair(dilation) != (1, 1):
        raise NotImplementedError("ic_gan_b200 conv2d_gradfix: dilation is not implemented (unused by StyleGAN2)")
    return _op(True, weight.shape, _square("stride", stride), _square("padding", padding),
               _square("output_padding", output_padding), groups).apply(input, weight, bias)
