"""Drop-in for ``stylegan2_ada_pytorch/torch_utils/ops/conv2d_resample.py`` (``conv2d_resample`` :79-216): 2-D convolution
with optional up/down-sampling, choosing between strided conv, transposed conv and FIR resampling exactly as the
reference does, on top of this package's ``conv2d_gradfix`` and ``upfirdn2d``."""
from __future__ import annotations

import torch

from . import conv2d_gradfix, upfirdn2d
from .upfirdn2d import _get_filter_size, _parse_padding


def _conv(x, w, stride=1, padding=0, groups=1, transpose=False, flip_weight=True):
    if not flip_weight:  # conv2d computes correlation; a true convolution needs the taps flipped
        w = w.flip([2, 3])
    op = conv2d_gradfix.conv_transpose2d if transpose else conv2d_gradfix.conv2d
    return op(x, w, stride=stride, padding=padding, groups=groups)


def conv2d_resample(x, w, f=None, up=1, down=1, padding=0, groups=1, flip_weight=True, flip_filter=False):
    assert isinstance(x, torch.Tensor) and x.ndim == 4 and isinstance(w, torch.Tensor) and w.ndim == 4
    assert isinstance(up, int) and up >= 1 and isinstance(down, int) and down >= 1 and groups >= 1
    co, ci_g, kh, kw = (int(s) for s in w.shape)
    fw, fh = _get_filter_size(f)
    px0, px1, py0, py1 = _parse_padding(padding)
    if up > 1:  # FIR padding that keeps the image aligned after zero-insertion
        px0 += (fw + up - 1) // 2
        px1 += (fw - up) // 2
        py0 += (fh + up - 1) // 2
        py1 += (fh - up) // 2
    if down > 1:
        px0 += (fw - down + 1) // 2
        px1 += (fw - down) // 2
        py0 += (fh - down + 1) // 2
        py1 += (fh - down) // 2
    pad4 = [px0, px1, py0, py1]

    if kw == 1 and kh == 1 and down > 1 and up == 1:  # 1x1 + downsample: filter/decimate first
        x = upfirdn2d.upfirdn2d(x, f, down=down, padding=pad4, flip_filter=flip_filter)
        return _conv(x, w, groups=groups, flip_weight=flip_weight)
    if kw == 1 and kh == 1 and up > 1 and down == 1:  # 1x1 + upsample: convolve at low resolution first
        x = _conv(x, w, groups=groups, flip_weight=flip_weight)
        return upfirdn2d.upfirdn2d(x, f, up=up, padding=pad4, gain=up ** 2, flip_filter=flip_filter)
    if down > 1 and up == 1:  # low-pass, then strided convolution
        x = upfirdn2d.upfirdn2d(x, f, padding=pad4, flip_filter=flip_filter)
        return _conv(x, w, stride=down, groups=groups, flip_weight=flip_weight)
    if up > 1:  # transposed (stride = up) convolution, then low-pass
        if groups == 1:
            wt = w.transpose(0, 1)
        else:
            wt = w.reshape(groups, co // groups, ci_g, kh, kw).transpose(1, 2).reshape(groups * ci_g, co // groups, kh, kw)
        px0 -= kw - 1
        px1 -= kw - up
        py0 -= kh - 1
        py1 -= kh - up
        pxt, pyt = max(min(-px0, -px1), 0), max(min(-py0, -py1), 0)
        x = _conv(x, wt, stride=up, padding=[pyt, pxt], groups=groups, transpose=True, flip_weight=(not flip_weight))
        x = upfirdn2d.upfirdn2d(x, f, padding=[px0 + pxt, px1 + pxt, py0 + pyt, py1 + pyt], gain=up ** 2,
                                flip_filter=flip_filter)
        if down > 1:
            x = upfirdn2d.upfirdn2d(x, f, down=down, flip_filter=flip_filter)
        return x
    if px0 == px1 and py0 == py1 and px0 >= 0 and py0 >= 0:  # plain convolution
        return _conv(x, w, padding=[py0, px0], groups=groups, flip_weight=flip_weight)
    x = upfirdn2d.upfirdn2d(x, (f if up > 1 else None), up=up, padding=pad4, gain=up ** 2, flip_filter=flip_filter)
    x = _conv(x, w, groups=groups, flip_weight=flip_weight)
    if down > 1:
        x = upfirdn2d.upfirdn2d(x, f, down=down, flip_filter=flip_filter)
    return x
