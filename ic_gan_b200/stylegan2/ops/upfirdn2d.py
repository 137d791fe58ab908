"""Drop-in for ``stylegan2_ada_pytorch/torch_utils/ops/upfirdn2d.py``: ``setup_filter`` (:88-139), ``upfirdn2d``
(:145-193), ``filter2d`` / ``upsample2d`` / ``downsample2d`` (:359-478), identical signatures.

On B200 every 4-D activation is channels-last, so the op runs on ``icgan_upfirdn2d_nhwc``: a shared-memory-tiled kernel
(one staged input patch per output tile, sliding register window, 128-bit accesses) for the 4x4 filters and the
(up, down) pairs StyleGAN2 uses; any other geometry -- separable 1-D filters, other factors, NCHW-contiguous tensors with
a channel count the vector path cannot take -- goes to the generic gather kernel ``icgan_upfirdn2d``.  The adjoint of
upfirdn2d is upfirdn2d with up/down swapped and the filter flipped (reference :324-349), so derivatives of any order
stay inside this one Function."""
from __future__ import annotations

import numpy as np
import torch

from ..._lib import call, dt, float_array, ptr, stream_ptr


# ------------------------------------------------------------------------------------------------- argument geometry
def _pair(v, what):
    a, b = (v, v) if isinstance(v, int) else tuple(v)
    a, b = int(a), int(b)
    if a < 1 or b < 1:
        raise ValueError(f"upfirdn2d: {what} factors must be >= 1")
    return a, b


def _pad4(padding):
    """int | (px, py) | (px0, px1, py0, py1) -> (px0, px1, py0, py1)."""
    if isinstance(padding, int):
        return (padding,) * 4
    p = [int(v) for v in padding]
    if len(p) == 2:
        return p[0], p[0], p[1], p[1]
    if len(p) != 4:
        raise ValueError("upfirdn2d: padding must have 1, 2 or 4 entries")
    return tuple(p)


def _taps(f):
    """(fw, fh) of a filter tensor; None is the identity filter."""
    if f is None:
        return 1, 1
    if not (isinstance(f, torch.Tensor) and f.ndim in (1, 2)):
        raise ValueError("upfirdn2d: filter must be a 1-D or 2-D tensor")
    return int(f.shape[-1]), int(f.shape[0])


# names the reference's other modules import from here
_parse_scaling = lambda scaling: _pair(scaling, "scaling")  # noqa: E731
_parse_padding = _pad4
_get_filter_size = _taps


def setup_filter(f, device=torch.device("cpu"), normalize=True, flip_filter=False, gain=1, separable=None):
    """FIR filter tensor in the form upfirdn2d expects (reference :88-139): 1-D taps become their outer product unless
    `separable` (default: only for >= 8 taps); normalised to unit sum; gain applied per axis."""
    f = torch.as_tensor(1 if f is None else f, dtype=torch.float32)
    if f.numel() == 0 or f.ndim > 2:
        raise ValueError("setup_filter: need a scalar, a vector or a matrix of taps")
    f = f.reshape(1) if f.ndim == 0 else f
    if separable is None:
        separable = f.ndim == 1 and f.numel() >= 8
    if f.ndim == 1 and not separable:
        f = torch.outer(f, f)
    if f.ndim != (1 if separable else 2):
        raise ValueError("setup_filter: a separable filter must be 1-D")
    if normalize:
        f = f / f.sum()
    if flip_filter:
        f = f.flip(list(range(f.ndim)))
    return (f * gain ** (f.ndim / 2)).to(device=device)


# ------------------------------------------------------------------------------------------------- launches
def _channels_last(x):
    return x.ndim == 4 and x.shape[1] > 1 and x.stride(1) == 1


def _fast_ok(x, f2d, up, down):
    nv = 4 if x.dtype == torch.float32 else 8
    return (tuple(f2d.shape) == (4, 4) and up[0] == up[1] and down[0] == down[1] and (up[0], down[0]) in ((1, 1), (2, 1), (1, 2))
            and x.shape[1] % nv == 0)


_factor_cache = {}


def _factors(f2d):
    """(fx, fy) host tuples if the 4x4 filter is an outer product fy (x) fx (every filter setup_filter builds from 1-D
    taps), else None.  One device->host read per filter buffer, cached on (pointer, version): it happens during the
    eager warm-up, never inside a captured graph."""
    key = (f2d.data_ptr(), f2d._version)
    if key not in _factor_cache:
        h = f2d.detach().double().cpu()
        total = float(h.sum())
        out = None
        if total != 0.0:
            fy, fx = h.sum(1) / total, h.sum(0)
            if float((torch.outer(fy, fx) - h).abs().max()) <= 1e-7 * float(h.abs().max()):
                out = (tuple(fx.tolist()), tuple(fy.tolist()))
        _factor_cache[key] = out
    return _factor_cache[key]


def _run(x, f2d, up, down, pad, flip, gain):
    N, C, H, W = x.shape
    fh, fw = f2d.shape
    ow = (W * up[0] + pad[0] + pad[1] - fw + down[0]) // down[0]
    oh = (H * up[1] + pad[2] + pad[3] - fh + down[1]) // down[1]
    f2d = f2d.contiguous()
    if _fast_ok(x, f2d, up, down):  # channels-last tiled kernel (a non-channels-last input is re-laid-out once)
        xin = x.contiguous(memory_format=torch.channels_last)
        y = torch.empty((N, C, oh, ow), device=x.device, dtype=x.dtype, memory_format=torch.channels_last)
        fac = _factors(f2d)
        call("icgan_upfirdn2d_nhwc", ptr(xin), ptr(f2d), ptr(y), N, C, H, W, up[0], down[0], pad[0], pad[1], pad[2], pad[3],
             int(flip), float(gain), None, None, None, 0, None, 0, 0.0, 1.0, -1.0, None, None,
             float_array(fac[0]) if fac else None, float_array(fac[1]) if fac else None, dt(x), stream_ptr())
        return y
    cl = _channels_last(x)
    xin = x.contiguous(memory_format=torch.channels_last) if cl else x.contiguous()
    y = torch.empty((N, C, oh, ow), device=x.device, dtype=x.dtype,
                    memory_format=torch.channels_last if cl else torch.contiguous_format)
    call("icgan_upfirdn2d", ptr(xin), ptr(f2d), ptr(y), N, C, H, W, fh, fw, up[0], up[1], down[0], down[1], pad[0], pad[1],
         pad[2], pad[3], int(flip), float(gain), int(cl), dt(x), stream_ptr())
    return y


_fn_cache = {}


def _function(up, down, pad, flip, gain):
    key = (up, down, pad, bool(flip), float(gain))
    if key in _fn_cache:
        return _fn_cache[key]

    class Upfirdn2dCuda(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, f):
            if f.ndim == 2:
                y = _run(x, f, up, down, pad, flip, gain)
            else:  # separable: one pass per axis, sqrt(gain) each (reference :281-283)
                y = _run(x, f.unsqueeze(0), (up[0], 1), (down[0], 1), (pad[0], pad[1], 0, 0), flip, np.sqrt(gain))
                y = _run(y, f.unsqueeze(1), (1, up[1]), (1, down[1]), (0, 0, pad[2], pad[3]), flip, np.sqrt(gain))
            ctx.save_for_backward(f)
            ctx.x_shape = x.shape
            return y

        @staticmethod
        def backward(ctx, dy):
            (f,) = ctx.saved_tensors
            if not ctx.needs_input_grad[0]:
                return None, None
            ih, iw = ctx.x_shape[2:]
            oh, ow = dy.shape[2:]
            fw, fh = _taps(f)
            adj = (fw - pad[0] - 1, iw * up[0] - ow * down[0] + pad[0] - up[0] + 1,
                   fh - pad[2] - 1, ih * up[1] - oh * down[1] + pad[2] - up[1] + 1)
            return _function(down, up, adj, not flip, gain).apply(dy, f), None

    _fn_cache[key] = Upfirdn2dCuda
    return Upfirdn2dCuda


def upfirdn2d(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1, impl="cuda"):
    if not (isinstance(x, torch.Tensor) and x.ndim == 4):
        raise ValueError("upfirdn2d: x must be a [N, C, H, W] tensor")
    if impl != "cuda":
        raise NotImplementedError("ic_gan_b200 has no PyTorch/CPU fallback for upfirdn2d (impl='ref' lives in oracle/)")
    if f is None:
        f = torch.ones([1, 1], dtype=torch.float32, device=x.device)
    f = f.to(device=x.device, dtype=torch.float32)
    return _function(_pair(up, "up"), _pair(down, "down"), _pad4(padding), flip_filter, gain).apply(x, f)


def _centred(padding, f, grow_x, grow_y):
    px0, px1, py0, py1 = _pad4(padding)
    return [px0 + grow_x[0], px1 + grow_x[1], py0 + grow_y[0], py1 + grow_y[1]]


def filter2d(x, f, padding=0, flip_filter=False, gain=1, impl="cuda"):
    """Same-size FIR filtering: the filter footprint is split around each sample (reference :359-392)."""
    fw, fh = _taps(f)
    p = _centred(padding, f, (fw // 2, (fw - 1) // 2), (fh // 2, (fh - 1) // 2))
    return upfirdn2d(x, f, padding=p, flip_filter=flip_filter, gain=gain, impl=impl)


def upsample2d(x, f, up=2, padding=0, flip_filter=False, gain=1, impl="cuda"):
    """Zero-stuff by `up`, filter, keep the image centred; gain up^2 restores the mean (reference :397-436)."""
    ux, uy = _pair(up, "up")
    fw, fh = _taps(f)
    p = _centred(padding, f, ((fw + ux - 1) // 2, (fw - ux) // 2), ((fh + uy - 1) // 2, (fh - uy) // 2))
    return upfirdn2d(x, f, up=up, padding=p, flip_filter=flip_filter, gain=gain * ux * uy, impl=impl)


def downsample2d(x, f, down=2, padding=0, flip_filter=False, gain=1, impl="cuda"):
    """Filter, then keep every `down`-th sample (reference :441-478)."""
    dx, dy = _pair(down, "down")
    fw, fh = _taps(f)
    p = _centred(padding, f, ((fw - dx + 1) // 2, (fw - dx) // 2), ((fh - dy + 1) // 2, (fh - dy) // 2))
    return upfirdn2d(x, f, down=down, padding=p, flip_filter=flip_filter, gain=gain, impl=impl)
