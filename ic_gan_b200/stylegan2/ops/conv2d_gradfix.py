"""Drop-in for ``stylegan2_ada_pytorch/torch_utils/ops/conv2d_gradfix.py``: ``conv2d`` (:43-62), ``conv_transpose2d``
(:65-94) and ``no_weight_gradients`` (:31-37), differentiable to arbitrary order (R1 and the path-length regulariser
take gradients of gradients through every convolution; reference :126-272).

B200 design.  Activations stay channels-last (NHWC in memory) from the first layer to the last, so the NCHW tensors the
callers see are free views.  Every convolution StyleGAN2 issues runs on the tcgen05 implicit-GEMM kernels of
libicgan_b200:

=====================================  ==================================================================================
3x3 pad 1 / 1x1, stride 1              ``icgan_conv2d_tc`` (halo-reuse kernel) -- forward, and dgrad with the flipped copy
3x3 stride 2 (D's down path)           ``icgan_conv2d_tc_ex`` with TMA traversal stride 2: no decimated copy of the input
3x3 stride-2 TRANSPOSED (G's up path,  four ``icgan_conv2d_tc_ex`` launches, one per output parity class with the 1/2/2/4
and the dgrad of the row above)        taps that reach it: 9 MACs per input pixel instead of 36 on a zero-stuffed tensor
weight gradients                       ``icgan_conv2d_wgrad_tc`` / ``_ex`` (MN-major operands straight from NHWC)
float32 tensors (non-fp16 blocks)      the same kernels on a split-bf16 representation x = hi + lo (hi.hi + lo.hi +
                                       hi.lo, fp32 accumulate: ~2^-16 relative, three launches)
<= 4 or odd channel counts (RGB side)  the streaming / CUDA-core kernels (``icgan_conv2d_small``, ``icgan_conv2d_simt``)
=====================================  ==================================================================================
"""
from __future__ import annotations

import contextlib

import torch

from ... import _lib as L
from ..._lib import call, dt, int_array, ptr, stream_ptr

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


def _one(name, v):
    a, b = (int(v), int(v)) if isinstance(v, int) else (int(v[0]), int(v[1]))
    if a != b:
        raise NotImplementedError(f"ic_gan_b200 conv2d_gradfix: anisotropic {name}={v} is not implemented")
    return a


# ------------------------------------------------------------------------------------------------- layout helpers
def nhwc(x: torch.Tensor) -> torch.Tensor:
    """[B,C,H,W] (any strides) -> contiguous [B,H,W,C] tensor sharing memory when x is already channels-last."""
    return x.permute(0, 2, 3, 1).contiguous()


def nchw_view(y: torch.Tensor) -> torch.Tensor:
    """contiguous [B,H,W,C] -> logical [B,C,H,W] (channels-last strides, no copy)."""
    return y.permute(0, 3, 1, 2)


def _work(x: torch.Tensor) -> torch.Tensor:
    return x.float() if x.dtype == torch.float16 else x  # half precision here is bfloat16; float16 is widened


def _split(t: torch.Tensor):
    """float32 -> (hi, lo) bfloat16 with hi + lo ~ t to 2^-16 relative."""
    hi = t.to(torch.bfloat16)
    return hi, (t - hi.float()).to(torch.bfloat16)


def _tc_ok(ci, co):
    return ci % 16 == 0 and co % 8 == 0


def _pad_last(t, mult):
    """Zero-pad the channel (last) dimension up to a multiple of `mult` (the 513-channel input of D's epilogue conv,
    networks.py:963: minibatch-std adds one channel; a few zero channels buy the tensor-core path)."""
    r = (-t.shape[-1]) % mult
    return t if r == 0 else torch.nn.functional.pad(t, (0, r))


# ------------------------------------------------------------------------------------------------- raw launches
def _timed(kind, flops, fn):
    """bench.py sets ops.PROFILE to a list: CUDA events around every tensor-core launch (the live roofline source)."""
    from ... import ops
    if ops.PROFILE is None:
        return fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    ops.PROFILE.append((kind, flops, e0, e1, None))


def _launch_same(x, wk, y, res, k):
    """stride 1, pad k//2 on the halo-reuse kernel; y (+)= conv(x, wk) where res (float32, may be y itself) is added."""
    B, H, W, ci = x.shape
    _timed("sg2_conv", 2.0 * B * H * W * ci * wk.shape[0] * k * k, lambda: call(
        "icgan_conv2d_tc", ptr(x), ptr(wk), None, None, ptr(res), ptr(y), None, B, H, W, ci, wk.shape[0], k, dt(y),
        dt(res) if res is not None else L.F32, 0, L.ACT_NONE, stream_ptr()))


def _launch_ex(x, wk, y, res, taps, in_stride, dom, omap):
    """taps: list of (dh, dw, weight slice); dom = (Hd, Wd); omap = (osy, ooy, osx, oox)."""
    B, H, W, ci = x.shape
    co, wt = wk.shape[0], wk.shape[1] * wk.shape[2]
    _timed("sg2_conv", 2.0 * B * dom[0] * dom[1] * ci * co * len(taps), lambda: call(
        "icgan_conv2d_tc_ex", ptr(x), ptr(wk), None, None, ptr(res), ptr(y), B, H, W, ci, co, wt, len(taps),
        int_array([t[0] for t in taps]), int_array([t[1] for t in taps]), int_array([t[2] for t in taps]), in_stride,
        dom[0], dom[1], y.shape[1], y.shape[2], omap[0], omap[1], omap[2], omap[3], dt(y),
        dt(res) if res is not None else L.F32, 0, stream_ptr()))


def _run_tc(x, wk32, launch, out_shape):
    """Run `launch(x_operand, w_operand, y, residual)` once (bf16 activations) or three times on split operands
    (float32 activations) and return y in x's dtype."""
    if x.dtype == torch.bfloat16:
        y = torch.empty(out_shape, device=x.device, dtype=torch.bfloat16)
        launch(x, wk32.to(torch.bfloat16), y, None)
        return y
    y = torch.empty(out_shape, device=x.device, dtype=torch.float32)
    xh, xl = _split(x)
    wh, wl = _split(wk32)
    launch(xh, wh, y, None)
    launch(xl, wh, y, y)
    launch(xh, wl, y, y)
    return y


# ------------------------------------------------------------------------------------------------- forward cores (NHWC)
def _conv_core(x, w, stride, pad):
    """x [B,H,W,Ci], w [Co,Ci,k,k] -> [B,Ho,Wo,Co]: y = corr(x, w)."""
    B, H, W, ci = x.shape
    co, k = w.shape[0], w.shape[2]
    ho, wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    wk = w.detach().float().permute(0, 2, 3, 1).contiguous()  # [Co,k,k,Ci]
    if not _tc_ok(ci, co) and min(ci, co) > 16 and x.dtype in (torch.bfloat16, torch.float32) and k * k <= 16:
        xp = _pad_last(x, 16)
        wp = _pad_last(wk, 16)
        if co % 8:
            wp = torch.nn.functional.pad(wp, (0, 0, 0, 0, 0, 0, 0, (-co) % 8))
        wfull = wp.permute(0, 3, 1, 2)  # back to [Co', Ci', k, k] for the recursive call
        return _conv_core(xp, wfull, stride, pad)[..., :co].contiguous()
    if _tc_ok(ci, co) and x.dtype in (torch.bfloat16, torch.float32) and k * k <= 16:
        if stride == 1 and pad == k // 2 and k in (1, 3):
            return _run_tc(x, wk, lambda a, b, y, r: _launch_same(a, b, y, r, k), (B, ho, wo, co))
        taps = [(kh - pad, kw - pad, kh * k + kw) for kh in range(k) for kw in range(k)]
        return _run_tc(x, wk, lambda a, b, y, r: _launch_ex(a, b, y, r, taps, stride, (ho, wo), (1, 0, 1, 0)),
                       (B, ho, wo, co))
    y = torch.empty(B, ho, wo, co, device=x.device, dtype=x.dtype)
    if min(ci, co) <= 4 and stride == 1 and pad == k // 2 and k in (1, 3):
        call("icgan_conv2d_small", ptr(x), ptr(wk), None, None, ptr(y), B, H, W, ci, co, k, dt(x), dt(y), L.ACT_NONE,
             stream_ptr())
    else:
        call("icgan_conv2d_simt", ptr(x), ptr(wk), None, None, None, ptr(y), B, H, W, ci, co, k, stride, pad, dt(x), dt(y),
             L.F32, 0, L.ACT_NONE, stream_ptr())
    return y


def _conv_transpose_core(x, w, stride, pad, out_pad):
    """x [B,H,W,Ci], w [Ci,Co,k,k] -> [B,Ho,Wo,Co], Ho = (H-1)*stride - 2*pad + k + out_pad:
    y[n, s*i + kh - pad, s*j + kw - pad, co] += x[n,i,j,ci] * w[ci,co,kh,kw]."""
    B, H, W, ci = x.shape
    co, k = w.shape[1], w.shape[2]
    ho, wo = (H - 1) * stride - 2 * pad + k + out_pad, (W - 1) * stride - 2 * pad + k + out_pad
    if stride == 1:  # a plain correlation with the taps reversed and the channel roles swapped
        return _conv_core_padded(x, w.flip([2, 3]).transpose(0, 1), k - 1 - pad, ho, wo)
    if stride != 2:
        raise NotImplementedError("ic_gan_b200 conv_transpose2d: stride 1 or 2")
    wk = w.detach().float().permute(1, 2, 3, 0).contiguous()  # [Co,k,k,Ci]
    if _tc_ok(ci, co) and x.dtype in (torch.bfloat16, torch.float32):
        def launch(a, b, y, r):
            for pa in (0, 1):          # output row parity: oy = 2*j + pa reads input row j + dh, dh = (pa + pad - kh) / 2
                for pb in (0, 1):
                    taps = [((pa + pad - kh) // 2, (pb + pad - kw) // 2, kh * k + kw)
                            for kh in range(k) if (pa + pad - kh) % 2 == 0
                            for kw in range(k) if (pb + pad - kw) % 2 == 0]
                    dom = ((ho - pa + 1) // 2, (wo - pb + 1) // 2)
                    if taps and dom[0] > 0 and dom[1] > 0:
                        _launch_ex(a, b, y, r, taps, 1, dom, (2, pa, 2, pb))
        return _run_tc(x, wk, launch, (B, ho, wo, co))
    # RGB-sized channel counts: zero-stuff + CUDA-core correlation (tiny layers only)
    q = k - 1 - pad
    one = torch.ones(1, 1, device=x.device, dtype=torch.float32)
    up = torch.empty(B, H * stride + 2 * q + out_pad - (stride - 1), W * stride + 2 * q + out_pad - (stride - 1), ci,
                     device=x.device, dtype=x.dtype)
    call("icgan_upfirdn2d", ptr(x), ptr(one), ptr(up), B, ci, H, W, 1, 1, stride, stride, 1, 1, q, q + out_pad - (stride - 1),
         q, q + out_pad - (stride - 1), 0, 1.0, 1, dt(x), stream_ptr())
    return _conv_core(up, w.flip([2, 3]).transpose(0, 1), 1, 0)


def _conv_core_padded(x, w, pad, ho, wo):
    y = _conv_core(x, w, 1, pad)
    assert y.shape[1] == ho and y.shape[2] == wo
    return y


def _wgrad_core(x, dy, stride, pad, k, transpose):
    """Weight gradient in the layout of the forward weight.  conv: dW[o,i,kh,kw] = sum dy[n,ho,wo,o] x[n,s*ho+kh-p,s*wo+kw-p,i];
    transposed conv (x [B,H,W,Ci] -> dy [B,Ho,Wo,Co], weight [Ci,Co,k,k]): dW[i,o,kh,kw] = sum x[n,h,w,i] dy[n,s*h+kh-p,s*w+kw-p,o]."""
    a, b = (dy, x) if not transpose else (x, dy)  # a indexes the result's rows and walks the tile domain unshifted
    B, ha, wa, ca = a.shape
    _, hb, wb, cb = b.shape
    if not _tc_ok(cb, ca) and min(ca, cb) > 16 and a.dtype in (torch.bfloat16, torch.float32) and k * k <= 16:
        ap, bp = _pad_last(a, 8), _pad_last(b, 16)
        xx, dd = (bp, ap) if not transpose else (ap, bp)
        return _wgrad_core(xx, dd, stride, pad, k, transpose)[:ca, :cb]
    if _tc_ok(cb, ca) and a.dtype in (torch.bfloat16, torch.float32) and k * k <= 16:
        G = torch.zeros(ca, k, k, cb, device=x.device, dtype=torch.float32)
        taps = [(kh - pad, kw - pad) for kh in range(k) for kw in range(k)]
        plain = stride == 1 and pad == k // 2 and k in (1, 3)

        def launch(aa, bb):
            flops = 2.0 * B * ha * wa * ca * cb * k * k
            if plain:
                _timed("sg2_wgrad", flops, lambda: call("icgan_conv2d_wgrad_tc", ptr(bb), ptr(aa), ptr(G), B, ha, wa, cb, ca,
                                                        k, stream_ptr()))
            else:
                _timed("sg2_wgrad", flops, lambda: call(
                    "icgan_conv2d_wgrad_tc_ex", ptr(aa), ptr(bb), ptr(G), B, ha, wa, ca, hb, wb, cb, k * k,
                    int_array([t[0] for t in taps]), int_array([t[1] for t in taps]), stride, stream_ptr()))
        if a.dtype == torch.bfloat16:
            launch(a, b.to(torch.bfloat16))
        else:
            (ah, al), (bh, bl) = _split(a), _split(b.float())
            launch(ah, bh); launch(al, bh); launch(ah, bl)
        return G.permute(0, 3, 1, 2)  # [ca, cb, k, k]: OIHW for conv, [Ci, Co, k, k] for the transposed form
    # CUDA-core path (RGB-sized layers); the transposed form swaps the operand roles of the plain one
    xin, gin = (x, dy) if not transpose else (dy, x)
    Bn, H, W, ci = xin.shape
    co = gin.shape[3]
    G = torch.zeros(co, k, k, ci, device=x.device, dtype=torch.float32)
    if gin.dtype != xin.dtype:
        gin = gin.to(xin.dtype)
    if min(ci, co) <= 4 and stride == 1 and pad == k // 2 and k in (1, 3):
        call("icgan_conv2d_wgrad_small", ptr(xin), ptr(gin), ptr(G), Bn, H, W, ci, co, k, dt(xin), dt(gin), stream_ptr())
    else:
        call("icgan_conv2d_wgrad_simt", ptr(xin), ptr(gin), ptr(G), Bn, H, W, ci, co, k, stride, pad, dt(xin), stream_ptr())
    return G.permute(0, 3, 1, 2)  # [Co,Ci,k,k]; for the transposed form the roles above already give [Ci,Co,k,k]


# ------------------------------------------------------------------------------------------------- autograd
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

    def run_one(x, w):
        xin = _work(nhwc(x))
        y = (_conv_transpose_core(xin, w, stride, padding, output_padding) if transpose
             else _conv_core(xin, w, stride, padding))
        return nchw_view(y).to(x.dtype)

    def run(x, w, b):
        if groups == 1:
            y = run_one(x, w)
        else:  # grouped form (only the fused-modconv inference path of the reference uses it; not on the training path)
            xs, ws = x.chunk(groups, 1), w.chunk(groups, 0)
            y = torch.cat([run_one(xi, wi) for xi, wi in zip(xs, ws)], 1)
        return y if b is None else y + b.to(y.dtype).reshape(1, -1, 1, 1)

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
                db = dy.sum([0, 2, 3], dtype=torch.float32).to(dy.dtype)
            return dx, dw, db

    class Conv2dGradWeight(torch.autograd.Function):
        @staticmethod
        def forward(ctx, dy, x):
            ctx.save_for_backward(dy, x)
            a, g = _work(nhwc(x)), _work(nhwc(dy))
            if g.dtype != a.dtype:
                g = g.to(a.dtype)
            if groups == 1:
                dw = _wgrad_core(a, g, stride, padding, k, transpose)
            else:
                dw = torch.cat([_wgrad_core(ai.contiguous(), gi.contiguous(), stride, padding, k, transpose)
                                for ai, gi in zip(a.chunk(groups, 3), g.chunk(groups, 3))], 0)
            return dw.to(x.dtype).contiguous()

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
    if _one("dilation", dilation) != 1:
        raise NotImplementedError("ic_gan_b200 conv2d_gradfix: dilation is not implemented (unused by StyleGAN2)")
    return _op(False, weight.shape, _one("stride", stride), _one("padding", padding), 0, groups).apply(input, weight, bias)


def conv_transpose2d(input, weight, bias=None, stride=1, padding=0, output_padding=0, groups=1, dilation=1):
    if _one("dilation", dilation) != 1:
        raise NotImplementedError("ic_gan_b200 conv2d_gradfix: dilation is not implemented (unused by StyleGAN2)")
    return _op(True, weight.shape, _one("stride", stride), _one("padding", padding), _one("output_padding", output_padding),
               groups).apply(input, weight, bias)
