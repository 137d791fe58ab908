"""Drop-in for ``stylegan2_ada_pytorch/torch_utils/ops/conv2d_resample.py`` (``conv2d_resample`` :79-216): convolution
with optional FIR up/down-sampling.

The function is  y = decimate_down( FIR_f( conv_w( pad( FIR_f*up^2( zero_stuff_up(x) ) ) ) ) )  and is always evaluated
as the cheapest equivalent chain of the two primitives this package runs natively -- ``upfirdn2d`` (shared-memory-tiled
NHWC kernel) and ``conv2d_gradfix`` (tcgen05 implicit GEMM, including stride 2 and the stride-2 transposed form).  The
chain is picked by :func:`plan`; it reproduces the evaluation order of the reference (which FIR runs before or after
which convolution decides the rounding, so parity needs the same order)."""
from __future__ import annotations

from typing import List, Tuple

import torch

from . import conv2d_gradfix, upfirdn2d

Step = Tuple  # ("fir", dict of upfirdn2d kwargs) | ("conv", dict(stride, padding, transpose))


def plan(kh: int, kw: int, fh: int, fw: int, up: int, down: int, padding) -> List[Step]:
    """Chain of primitive steps for one geometry (pure arithmetic, unit-tested on its own)."""
    left, right, top, bottom = upfirdn2d._pad4(padding)
    if up > 1:      # keep the image centred under the interpolation filter
        left, right = left + (fw + up - 1) // 2, right + (fw - up) // 2
        top, bottom = top + (fh + up - 1) // 2, bottom + (fh - up) // 2
    if down > 1:    # ... and under the anti-aliasing filter
        left, right = left + (fw - down + 1) // 2, right + (fw - down) // 2
        top, bottom = top + (fh - down + 1) // 2, bottom + (fh - down) // 2
    frame = [left, right, top, bottom]
    pointwise = kh == 1 and kw == 1

    if pointwise and up == 1 and down > 1:       # decimate first: the 1x1 convolution then sees down^2 fewer pixels
        return [("fir", dict(down=down, padding=frame)), ("conv", dict())]
    if pointwise and down == 1 and up > 1:       # 1x1 at low resolution, interpolate afterwards
        return [("conv", dict()), ("fir", dict(up=up, padding=frame, gain=up ** 2))]
    if up == 1 and down > 1:                     # low-pass at full resolution, stride-`down` convolution
        return [("fir", dict(padding=frame)), ("conv", dict(stride=down))]
    if up > 1:                                   # stride-`up` transposed convolution, then the interpolation filter
        # the transposed convolution already spreads each sample over a kh x kw footprint: take that out of the frame
        left, right = left - (kw - 1), right - (kw - up)
        top, bottom = top - (kh - 1), bottom - (kh - up)
        crop_x, crop_y = max(min(-left, -right), 0), max(min(-top, -bottom), 0)
        steps = [("conv", dict(stride=up, padding=[crop_y, crop_x], transpose=True)),
                 ("fir", dict(padding=[left + crop_x, right + crop_x, top + crop_y, bottom + crop_y], gain=up ** 2))]
        if down > 1:
            steps.append(("fir", dict(down=down)))
        return steps
    if left == right and top == bottom and left >= 0 and top >= 0:   # nothing to resample
        return [("conv", dict(padding=[top, left]))]
    steps = [("fir", dict(up=up, padding=frame, gain=up ** 2, no_filter=(up == 1))), ("conv", dict())]
    if down > 1:
        steps.append(("fir", dict(down=down)))
    return steps


def conv2d_resample(x, w, f=None, up=1, down=1, padding=0, groups=1, flip_weight=True, flip_filter=False):
    if not (isinstance(x, torch.Tensor) and x.ndim == 4 and isinstance(w, torch.Tensor) and w.ndim == 4):
        raise ValueError("conv2d_resample: x and w must be 4-D tensors")
    if not (isinstance(up, int) and up >= 1 and isinstance(down, int) and down >= 1 and groups >= 1):
        raise ValueError("conv2d_resample: up, down, groups must be positive integers")
    co, ci_g, kh, kw = (int(v) for v in w.shape)
    fw, fh = upfirdn2d._taps(f)
    for kind, kw_args in plan(kh, kw, fh, fw, up, down, padding):
        if kind == "fir":
            args = dict(kw_args)
            filt = None if args.pop("no_filter", False) else f
            x = upfirdn2d.upfirdn2d(x, filt, flip_filter=flip_filter, **args)
            continue
        transpose = kw_args.get("transpose", False)
        weight = w
        if transpose:  # [Co, Ci/g, kh, kw] -> the [Ci, Co/g, kh, kw] layout of a transposed convolution
            weight = (w.transpose(0, 1) if groups == 1 else
                      w.reshape(groups, co // groups, ci_g, kh, kw).transpose(1, 2).reshape(groups * ci_g, co // groups, kh, kw))
        # conv2d computes a correlation; a true convolution (flip_weight=False) reverses the taps -- and a transposed
        # convolution reverses them once more
        if flip_weight == transpose:
            weight = weight.flip([2, 3])
        op = conv2d_gradfix.conv_transpose2d if transpose else conv2d_gradfix.conv2d
        x = op(x, weight, stride=kw_args.get("stride", 1), padding=kw_args.get("padding", 0), groups=groups)
    return x
