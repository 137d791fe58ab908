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


def conv2d_resample(x, w, f=None, up=1, down=1, padding=0, flip_weight=True):
    """The reference's generic fallback path (conv2d_resample.py:204-216): FIR-upsample, convolve, FIR-downsample. Its
    specialised fast paths (strided / transposed convolutions) are algebraically identical to this definition."""
    fw = fh = 1 if f is None else int(f.shape[-1])
    if isinstance(padding, int):
        padding = [padding] * 4
    if len(padding) == 2:
        padding = [padding[0], padding[0], padding[1], padding[1]]
    px0, px1, py0, py1 = padding
    if up > 1:
        px0 += (fw + up - 1) // 2
        px1 += (fw - up) // 2
        py0 += (fh + up - 1) // 2
        py1 += (fh - up) // 2
    if down > 1:
        px0 += (fw - down + 1) // 2
        px1 += (fw - down) // 2
        py0 += (fh - down + 1) // 2
        py1 += (fh - down) // 2
    x = upfirdn2d(x, f if up > 1 else None, up=up, padding=[px0, px1, py0, py1], gain=up ** 2)
    x = F.conv2d(x, w if flip_weight else w.flip([2, 3]))
    if down > 1:
        x = upfirdn2d(x, f, down=down)
    return x


def modulated_conv2d(x, weight, styles, noise=None, up=1, down=1, padding=0, resample_filter=None, demodulate=True):
    """networks.py:37-117 in its mathematical form: per-sample weights w_n = W * s_n, demodulated, one conv per sample."""
    outs = []
    for n in range(x.shape[0]):
        w = weight * styles[n].reshape(1, -1, 1, 1)
        if demodulate:
            w = w * (w.square().sum(dim=[1, 2, 3], keepdim=True) + 1e-8).rsqrt()
        outs.append(conv2d_resample(x[n:n + 1], w, resample_filter, up=up, down=down, padding=padding))
    y = torch.cat(outs, 0)
    return y if noise is None else y + noise


CONV_SITES = [  # (name, Cin, Cout, k, up, down, padding, H)
    ("plain3", 8, 6, 3, 1, 1, 1, 8),
    ("up3", 8, 6, 3, 2, 1, 1, 8),
    ("down3", 8, 6, 3, 1, 2, 1, 8),
    ("down1_skip", 8, 6, 1, 1, 2, 0, 8),
    ("torgb1", 8, 3, 1, 1, 1, 0, 8),
    ("up1", 8, 6, 1, 2, 1, 0, 8),
]
