"""``fma(a, b, c) = a * b + c`` (stylegan2_ada_pytorch/torch_utils/ops/fma.py:19-52).

In the reference this exists so that ``x * dcoefs + noise`` (networks.py:88-89) costs one pass and its broadcast
gradients are reduced explicitly.  On B200 that expression never appears on its own -- the demodulation scale and the
noise ride in the activation kernel (``elementwise.mod_bias_act``) -- but the op keeps the reference's name and
signature for callers outside the hot path: the [N,C,H,W] * [N,C,1,1] + [N,1,H,W] pattern runs on the same fused kernel
(act = linear), any other broadcasting pattern is a plain expression."""
from __future__ import annotations

import torch


def fma(a, b, c):
    from . import elementwise
    if (a.ndim == 4 and a.is_cuda and b.ndim == 4 and c.ndim == 4 and b.shape[2:] == (1, 1) and b.shape[:2] == a.shape[:2]
            and c.shape[1] == 1 and c.shape[2:] == a.shape[2:] and c.shape[0] in (1, a.shape[0]) and a.shape[1] % 8 == 0
            and a.dtype in (torch.float32, torch.bfloat16)):
        return elementwise.mod_bias_act(a, pre=b.reshape(b.shape[0], -1), noise=c, bias=None, act="linear", gain=1.0, clamp=None)
    return a * b + c
