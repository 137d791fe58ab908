"""BigGAN-deep residual blocks on the B200 kernels.

Same constructor keywords, attribute names (``conv1..conv4``, ``bn1..bn4``, ``conv_sc``) and forward signatures as
``BigGAN_PyTorch/BigGANdeep.py`` ``GBlock`` (:33-85) and ``DBlock`` (:394-451): bottleneck blocks (1x1 down to
``channels // channel_ratio``, two 3x3, 1x1 back up) whose shortcut carries no weights in G (channels are dropped, not
projected) and concatenates ``conv_sc`` outputs in D.  Built from the same ops as the plain BigGAN blocks
(:mod:`ic_gan_b200.biggan.layers`): the nearest-neighbour upsampling of the shortcut is read at half resolution inside
conv4's epilogue, batch-norm + ReLU (+ upsampling) is one pass, every ReLU of the D block rides in the producing
convolution's epilogue.  (No IC-GAN config selects the deep generator -- BigGANdeep.py has no instance conditioning --
so only the blocks, which BASELINE's north_star names, are provided.)"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from .._lib import ACT_RELU
from . import layers
from .layers import to_nchw, to_nhwc


def _bn_relu(bn_mod, x, y, up=False):
    if isinstance(bn_mod, layers.ccbn):
        return bn_mod.fused(x, y, relu=True, up=up)
    return ops.BNActFn.apply(x, bn_mod.gain, bn_mod.bias, bn_mod.stored_mean, bn_mod.stored_var, bn_mod.training,
                             bn_mod.eps, bn_mod.momentum, True, up, x.dtype, None, None, bn_mod._stat_hint)


class GBlock(nn.Module):
    def __init__(self, in_channels, out_channels, which_conv=nn.Conv2d, which_bn=layers.bn, activation=None,
                 upsample=None, channel_ratio=4):
        super().__init__()
        if not layers._is_relu(activation):
            raise NotImplementedError("BigGAN-deep GBlock on B200 fuses ReLU; other activations are not implemented")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.hidden_channels = in_channels // channel_ratio
        self.which_conv, self.which_bn, self.activation, self.upsample = which_conv, which_bn, activation, upsample
        self.conv1 = which_conv(in_channels, self.hidden_channels, kernel_size=1, padding=0)
        self.conv2 = which_conv(self.hidden_channels, self.hidden_channels)
        self.conv3 = which_conv(self.hidden_channels, self.hidden_channels)
        self.conv4 = which_conv(self.hidden_channels, out_channels, kernel_size=1, padding=0)
        self.bn1 = which_bn(in_channels)
        self.bn2 = which_bn(self.hidden_channels)
        self.bn3 = which_bn(self.hidden_channels)
        self.bn4 = which_bn(self.hidden_channels)

    def forward_nhwc(self, x, y):
        up = bool(self.upsample)
        h = self.conv1.conv_nhwc(_bn_relu(self.bn1, x, y))
        h = _bn_relu(self.bn2, h, y, up=up)                       # BN-ReLU and the x2 upsampling in one pass
        sc = x if self.in_channels == self.out_channels else x[..., :self.out_channels].contiguous()
        h = self.conv2.conv_nhwc(h)
        h = self.conv3.conv_nhwc(_bn_relu(self.bn3, h, y))
        h = _bn_relu(self.bn4, h, y)
        return self.conv4.conv_nhwc(h, residual=sc, res_shift=1 if up else 0)  # + (upsampled) shortcut in the epilogue

    def forward(self, x, y):
        return to_nchw(self.forward_nhwc(to_nhwc(x), y))


class DBlock(nn.Module):
    def __init__(self, in_channels, out_channels, which_conv=layers.SNConv2d, wide=True, preactivation=True,
                 activation=None, downsample=None, channel_ratio=4):
        super().__init__()
        if not layers._is_relu(activation):
            raise NotImplementedError("BigGAN-deep DBlock on B200 fuses ReLU; other activations are not implemented")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.hidden_channels = out_channels // channel_ratio
        self.which_conv, self.preactivation, self.activation, self.downsample = which_conv, preactivation, activation, downsample
        self.conv1 = which_conv(in_channels, self.hidden_channels, kernel_size=1, padding=0)
        self.conv2 = which_conv(self.hidden_channels, self.hidden_channels)
        self.conv3 = which_conv(self.hidden_channels, self.hidden_channels)
        self.conv4 = which_conv(self.hidden_channels, out_channels, kernel_size=1, padding=0)
        self.learnable_sc = in_channels != out_channels
        if self.learnable_sc:
            self.conv_sc = which_conv(in_channels, out_channels - in_channels, kernel_size=1, padding=0)

    def forward_nhwc(self, x):
        down = bool(self.downsample)
        h = self.conv1.conv_nhwc(ops.ReluFn.apply(x), act=ACT_RELU)   # conv1(relu(x)), then the ReLU before conv2
        h = self.conv2.conv_nhwc(h, act=ACT_RELU)
        h = self.conv3.conv_nhwc(h, act=ACT_RELU)                     # "relu before downsample"
        if down:
            h = ops.Pool2Fn.apply(h, None, 0.25, 0)
        s = ops.Pool2Fn.apply(x, None, 0.25, 0) if down else x
        if self.learnable_sc:
            s = torch.cat([s, self.conv_sc.conv_nhwc(s)], dim=3)
        return self.conv4.conv_nhwc(h, residual=s)

    def forward(self, x):
        xin = to_nhwc(x)
        if xin.dtype != self.conv1.compute_dtype:
            xin = xin.to(self.conv1.compute_dtype)
        return to_nchw(self.forward_nhwc(xin))
