"""``modulated_conv2d`` of StyleGAN2 (stylegan2_ada_pytorch/training/networks.py:37-117), B200 formulation.

The reference has two evaluation orders: training scales the ACTIVATIONS (``x*styles -> conv(W) -> *dcoefs + noise``,
:77-95), inference scales the WEIGHTS per sample and runs one grouped convolution (:98-117).  They are the same function
of (x, W, styles, noise).  Here there is one order -- the activation-scaling one, for both -- because on B200 the
per-sample weights are the expensive object (B x Cout x Cin x k^2 operands instead of one shared tensor-core operand):

* demodulation coefficients without the [B,Co,Ci,k,k] temporary:  dcoefs = rsqrt(styles^2 @ (sum_k W^2)^T + 1e-8);
* ``x * styles`` is one channels-last pass (``elementwise.modulate``) whose output feeds the tensor-core convolution;
* ``* dcoefs + noise`` -- and, through :func:`modulated_conv2d_act`, the layer's bias, leaky ReLU, gain and clamp
  (networks.py:441-444) -- is ONE pass (``elementwise.mod_bias_act``) instead of ``fma`` followed by ``bias_act``.

``fused_modconv`` is accepted and ignored."""
from __future__ import annotations

import torch

from .ops import conv2d_resample, elementwise


def demod_coefficients(weight, styles):
    """[N, Cout] float32: rsqrt(sum_{ci,k} (W[co,ci,k] * s[n,ci])^2 + 1e-8)."""
    w2 = weight.float().square().sum(dim=[2, 3])            # [Co, Ci]
    return (styles.float().square() @ w2.t() + 1e-8).rsqrt()


def modulated_conv2d_act(x, weight, styles, noise=None, bias=None, act="linear", gain=1.0, clamp=None, up=1, down=1,
                         padding=0, resample_filter=None, demodulate=True, flip_weight=True):
    """bias_act(modulated_conv2d(x, ...), bias, act, gain, clamp) with the epilogue in a single pass."""
    n = int(x.shape[0])
    styles = styles.float().reshape(n, -1)
    dcoefs = demod_coefficients(weight, styles) if demodulate else None
    xs = elementwise.modulate(x, styles)
    y = conv2d_resample.conv2d_resample(x=xs, w=weight.to(x.dtype), f=resample_filter, up=up, down=down,
                                        padding=padding, flip_weight=flip_weight)
    if dcoefs is None and noise is None and bias is None and act == "linear" and gain == 1 and clamp is None:
        return y
    return elementwise.mod_bias_act(y, pre=dcoefs, noise=noise, bias=bias, act=act, gain=gain, clamp=clamp)


def modulated_conv2d(x, weight, styles, noise=None, up=1, down=1, padding=0, resample_filter=None, demodulate=True,
                     flip_weight=True, fused_modconv=True):
    """Reference signature (networks.py:37-50).  `noise` broadcasts against [N, Cout, H, W]."""
    if noise is not None and noise.ndim < 4:
        noise = noise.reshape((1,) * (4 - noise.ndim) + tuple(noise.shape))
    return modulated_conv2d_act(x, weight, styles, noise=noise, up=up, down=down, padding=padding,
                                resample_filter=resample_filter, demodulate=demodulate, flip_weight=flip_weight)
