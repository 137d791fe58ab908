"""CPU oracle for StyleGAN2-ADA's two native ops — TEST INFRASTRUCTURE ONLY (never imported by ic_gan_b200/).

Restates the reference's own slow implementations: ``_bias_act_ref`` (torch_utils/ops/bias_act.py:178-207, activation
table :26-99) and ``_upfirdn2d_ref`` (torch_utils/ops/upfirdn2d.py:200-246) in plain float32 PyTorch; gradients of any
order come from autograd.  Pinned against the live reference (impl='ref') by oracle/make_golden_extra.py.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

ACTS = {  # name -> (fn(x, alpha), def_alpha, def_gain)
    "linear": (lambda x, a: x, 0.0, 1.0),
    "relu": (lambda x, a: torch.relu(x), 0.0, math.sqrt(2)),
    "lrelu": (lambda x, a: F.leaky_relu(x, a), 0.2, math.sqrt(2)),
    "tanh": (lambda x, a: torch.tanh(x), 0.0, 1.0),
    "sigmoid": (lambda x, a: torch.sigmoid(x), 0.0, 1.0),
    "elu": (lambda x, a: F.elu(x), 0.0, 1.0),
    "selu": (lambda x, a: F.selu(x), 0.0, 1.0),
    "softplus": (lambda x, a: F.softplus(x), 0.0, 1.0),
    "swish": (lambda x, a: torch.sigmoid(x) * x, 0.0, math.sqrt(2)),
}


def bias_act(x, b=None, dim=1, act="linear", alpha=None, gain=None, clamp=None):
    fn, def_alpha, def_gain = ACTS[act]
    alpha = float(alpha if alpha is not None else def_alpha)
    gain = float(gain if gain is not None else def_gain)
    clamp = float(clamp if clamp is not None else -1)
    if b is not None:
        x = x + b.reshape([-1 if i == dim else 1 for i in range(x.ndim)])
    x = fn(x, alpha)
    if gain != 1:
        x = x * gain
    if clamp >= 0:
        x = x.clamp(-clamp, clamp)
    return x


def upfirdn2d(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1):
    upx, upy = (up, up) if isinstance(up, int) else up
    downx, downy = (down, down) if isinstance(down, int) else down
    if isinstance(padding, int):
        padding = [padding] * 4
    if len(padding) == 2:
        padding = [padding[0], padding[0], padding[1], padding[1]]
    padx0, padx1, pady0, pady1 = padding
    n, c, h, w = x.shape
    if f is None:
        f = torch.ones([1, 1], dtype=torch.float32)
    # zero-insertion upsampling
    x = x.reshape(n, c, h, 1, w, 1)
    x = F.pad(x, [0, upx - 1, 0, 0, 0, upy - 1])
    x = x.reshape(n, c, h * upy, w * upx)
    # pad (negative = crop)
    x = F.pad(x, [max(padx0, 0), max(padx1, 0), max(pady0, 0), max(pady1, 0)])
    x = x[:, :, max(-pady0, 0): x.shape[2] - max(-pady1, 0), max(-padx0, 0): x.shape[3] - max(-padx1, 0)]
    f = f * (gain ** (f.ndim / 2))
    f = f.to(x.dtype)
    if not flip_filter:
        f = f.flip(list(range(f.ndim)))
    f = f[None, None].repeat([c, 1] + [1] * f.ndim)
    if f.ndim == 4:
        x = F.conv2d(x, f, groups=c)
    else:
        x = F.conv2d(x, f.unsqueeze(2), groups=c)
        x = F.conv2d(x, f.unsqueeze(3), groups=c)
    return x[:, :, ::downy, ::downx]


def setup_filter(taps, gain=1.0):
    """2-D outer-product FIR normalised to sum 1 (upfirdn2d.setup_filter for 1-D taps with < 8 entries)."""
    f = torch.as_tensor(taps, dtype=torch.float32)
    f = f.ger(f)
    f = f / f.sum()
    return f * gain


# the upfirdn2d call sites of the IC-GAN StyleGAN2 256^2 networks
UPFIRDN_SITES = [
    dict(name="G_after_convT", up=1, down=1, padding=[1, 1, 1, 1], gain=4.0, odd=True),
    dict(name="G_rgb_upsample", up=2, down=1, padding=[2, 1, 2, 1], gain=4.0, odd=False),
    dict(name="D_skip_down", up=1, down=2, padding=[1, 1, 1, 1], gain=1.0, odd=False),
    dict(name="D_before_stride2_conv", up=1, down=1, padding=[2, 2, 2, 2], gain=1.0, odd=False),
]
