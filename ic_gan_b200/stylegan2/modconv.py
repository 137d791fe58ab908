"""``modulated_conv2d`` of StyleGAN2 (stylegan2_ada_pytorch/training/networks.py:37-117) on the B200 ops.

Training uses the non-fused form (networks.py:77-95): scale the input by the per-sample styles, convolve with the shared
weight through ``conv2d_resample``, scale the output by the demodulation coefficients and add noise (``fma``).  Inference
uses the fused form (:98-117): per-sample weights as one grouped convolution."""
from __future__ import annotations

import numpy as np
import torch

from .ops import conv2d_resample, fma


def modulated_conv2d(x, weight, styles, noise=None, up=1, down=1, padding=0, resample_filter=None, demodulate=True,
                     flip_weight=True, fused_modconv=True):
    batch = int(x.shape[0])
    co, ci, kh, kw = weight.shape
    if x.dtype == torch.float16 and demodulate:  # pre-normalise to avoid fp16 overflow (networks.py:57-63)
        weight = weight * (1 / np.sqrt(ci * kh * kw) / weight.norm(float("inf"), dim=[1, 2, 3], keepdim=True))
        styles = styles / styles.norm(float("inf"), dim=1, keepdim=True)
    w = dcoefs = None
    if demodulate or fused_modconv:
        w = weight.unsqueeze(0) * styles.reshape(batch, 1, -1, 1, 1)
    if demodulate:
        dcoefs = (w.square().sum(dim=[2, 3, 4]) + 1e-8).rsqrt()
    if demodulate and fused_modconv:
        w = w * dcoefs.reshape(batch, -1, 1, 1, 1)
    if not fused_modconv:
        x = x * styles.to(x.dtype).reshape(batch, -1, 1, 1)
        x = conv2d_resample.conv2d_resample(x=x, w=weight.to(x.dtype), f=resample_filter, up=up, down=down,
                                            padding=padding, flip_weight=flip_weight)
        if demodulate and noise is not None:
            x = fma.fma(x, dcoefs.to(x.dtype).reshape(batch, -1, 1, 1), noise.to(x.dtype))
        elif demodulate:
            x = x * dcoefs.to(x.dtype).reshape(batch, -1, 1, 1)
        elif noise is not None:
            x = x.add_(noise.to(x.dtype))
        return x
    x = x.reshape(1, -1, *x.shape[2:])
    w = w.reshape(-1, ci, kh, kw)
    x = conv2d_resample.conv2d_resample(x=x, w=w.to(x.dtype), f=resample_filter, up=up, down=down, padding=padding,
                                        groups=batch, flip_weight=flip_weight)
    x = x.reshape(batch, -1, *x.shape[2:])
    if noise is not None:
        x = x.add_(noise)
    return x
