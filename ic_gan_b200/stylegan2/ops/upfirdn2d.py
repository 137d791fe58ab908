"""Drop-in for ``stylegan2_ada_pytorch/torch_utils/ops/upfirdn2d.py``: ``setup_filter`` (:88-139), ``upfirdn2d``
(:145-193), ``filter2d`` / ``upsample2d`` / ``downsample2d`` (:359-478) with identical signatures, on ``icgan_upfirdn2d``.
The backward of upfirdn2d is upfirdn2d itself with up/down swapped and the filter flipped (:324-349), so arbitrary-order
derivatives (R1, path length) work."""
from __future__ import annotations

import numpy as np
import torch

from ..._lib import call, dt, ptr, stream_ptr


def _parse_scaling(scaling):
    if isinstance(scaling, int):
        scaling = [scaling, scaling]
    sx, sy = scaling
    assert sx >= 1 and sy >= 1
    return int(sx), int(sy)


def _parse_padding(padding):
    if isinstance(padding, int):
        padding = [padding, padding]
    if len(padding) == 2:
        px, py = padding
        padding = [px, px, py, py]
    return tuple(int(v) for v in padding)


def _get_filter_size(f):
    if f is None:
        return 1, 1
    assert isinstance(f, torch.Tensor) and f.ndim in [1, 2]
    return int(f.shape[-1]), int(f.shape[0])


def setup_filter(f, device=torch.device("cpu"), normalize=True, flip_filter=False, gain=1, separable=None):
    if f is None:
        f = 1
    f = torch.as_tensor(f, dtype=torch.float32)
    assert f.ndim in [0, 1, 2] and f.numel() > 0
    if f.ndim == 0:
        f = f[np.newaxis]
    if separable is None:
        separable = f.ndim == 1 and f.numel() >= 8
    if f.ndim == 1 and not separable:
        f = f.ger(f)
    assert f.ndim == (1 if separable else 2)
    if normalize:
        f = f / f.sum()
    if flip_filter:
        f = f.flip(list(range(f.ndim)))
    f = f * (gain ** (f.ndim / 2))
    return f.to(device=device)


def _run(x, f2d, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain):
    cl = x.ndim == 4 and x.stride(1) == 1 and x.shape[1] > 1
    x = x.contiguous(memory_format=torch.channels_last) if cl else x.contiguous()
    N, C, H, W = x.shape
    fh, fw = f2d.shape
    ow = (W * upx + padx0 + padx1 - fw + downx) // downx
    oh = (H * upy + pady0 + pady1 - fh + downy) // downy
    y = torch.empty((N, C, oh, ow), device=x.device, dtype=x.dtype,
                    memory_format=torch.channels_last if cl else torch.contiguous_format)
    f2d = f2d.contiguous()
    call("icgan_upfirdn2d", ptr(x), ptr(f2d), ptr(y), N, C, H, W, fh, fw, upx, upy, downx, downy, padx0,
         padx1, pady0, pady1, int(flip), float(gain), int(cl), dt(x), stream_ptr())
    return y


def upfirdn2d(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1, impl="cuda"):
    assert isinstance(x, torch.Tensor) and x.ndim == 4
    if impl != "cuda":
        raise NotImplementedError("ic_gan_b200 has no PyTorch/CPU fallback for upfirdn2d (impl='ref' lives in oracle/)")
    upx, upy = _parse_scaling(up)
    downx, downy = _parse_scaling(down)
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    if f is None:
        f = torch.ones([1, 1], dtype=torch.float32, device=x.device)
    f = f.to(device=x.device, dtype=torch.float32)

    class Upfirdn2dCuda(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, f):
            if f.ndim == 2:
                y = _run(x, f, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip_filter, gain)
            else:  # separable 1-D filter: two passes, gain split as sqrt(gain) each (upfirdn2d.py:281-283)
                y = _run(x, f.unsqueeze(0), upx, 1, downx, 1, padx0, padx1, 0, 0, flip_filter, np.sqrt(gain))
                y = _run(y, f.unsqueeze(1), 1, upy, 1, downy, 0, 0, pady0, pady1, flip_filter, np.sqrt(gain))
            ctx.save_for_backward(f)
            ctx.x_shape = x.shape
            return y

        @staticmethod
        def backward(ctx, dy):
            (f,) = ctx.saved_tensors
            _, _, ih, iw = ctx.x_shape
            _, _, oh, ow = dy.shape
            fw, fh = _get_filter_size(f)
            p = [fw - padx0 - 1, iw * upx - ow * downx + padx0 - upx + 1,
                 fh - pady0 - 1, ih * upy - oh * downy + pady0 - upy + 1]
            dx = None
            if ctx.needs_input_grad[0]:
                dx = upfirdn2d(dy, f, up=[downx, downy], down=[upx, upy], padding=p, flip_filter=(not flip_filter),
                               gain=gain)
            return dx, None

    return Upfirdn2dCuda.apply(x, f)


def filter2d(x, f, padding=0, flip_filter=False, gain=1, impl="cuda"):
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    p = [padx0 + fw // 2, padx1 + (fw - 1) // 2, pady0 + fh // 2, pady1 + (fh - 1) // 2]
    return upfirdn2d(x, f, padding=p, flip_filter=flip_filter, gain=gain, impl=impl)


def upsample2d(x, f, up=2, padding=0, flip_filter=False, gain=1, impl="cuda"):
    upx, upy = _parse_scaling(up)
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    p = [padx0 + (fw + upx - 1) // 2, padx1 + (fw - upx) // 2, pady0 + (fh + upy - 1) // 2, pady1 + (fh - upy) // 2]
    return upfirdn2d(x, f, up=up, padding=p, flip_filter=flip_filter, gain=gain * upx * upy, impl=impl)


def downsample2d(x, f, down=2, padding=0, flip_filter=False, gain=1, impl="cuda"):
    downx, downy = _parse_scaling(down)
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    p = [padx0 + (fw - downx + 1) // 2, padx1 + (fw - downx) // 2, pady0 + (fh - downy + 1) // 2,
         pady1 + (fh - downy) // 2]
    return upfirdn2d(x, f, down=down, padding=p, flip_filter=flip_filter, gain=gain, impl=impl)
