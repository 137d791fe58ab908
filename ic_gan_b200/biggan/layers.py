"""Drop-in ``layers`` module for IC-GAN's BigGAN backbone on B200.

Same class names, constructor keywords, ``forward`` signatures and ``state_dict`` keys/shapes as
``BigGAN_PyTorch/layers.py`` of facebookresearch/ic_gan (SNConv2d :116-153, SNLinear :157-165, SNEmbedding :171-200,
Attention :206-244, ccbn :359-442, bn :446-503, GBlock :512-552, DBlock :556-613), so reference checkpoints load with
``strict=True`` and ``isinstance(m, nn.Conv2d / nn.Linear / nn.Embedding)`` based initialisers keep working.  The math
runs in the hand-written sm_100a kernels of libicgan_b200 through :mod:`ic_gan_b200.ops`; there is no PyTorch/cuDNN
fallback.

Tensors crossing a module boundary are logical NCHW (as in the reference) but physically channels-last; inside the
blocks everything is explicit NHWC.
"""
from __future__ import annotations

import functools

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn import Parameter as P

from .. import ops
from .._lib import ACT_NONE, ACT_RELU, ACT_TANH

__all__ = ["SN", "SNConv2d", "SNLinear", "SNEmbedding", "Attention", "ccbn", "bn", "GBlock", "DBlock", "identity"]


def to_nhwc(x: torch.Tensor) -> torch.Tensor:
    """logical NCHW -> contiguous [B,H,W,C] (free when x is already channels-last)."""
    return x.permute(0, 2, 3, 1).contiguous()


def to_nchw(x: torch.Tensor) -> torch.Tensor:
    return x.permute(0, 3, 1, 2)


class identity(nn.Module):
    def forward(self, input):
        return input


def _is_relu(fn) -> bool:
    return fn is None or isinstance(fn, nn.ReLU) or fn is F.relu or fn is torch.relu


# --------------------------------------------------------------------------------------------- spectral norm
class SN(object):
    """Spectral-norm mix-in: registers the reference's ``u0`` / ``sv0`` buffers and owns the device-side SN state."""

    compute_dtype = torch.float32

    def _sn_init(self, kind, num_svs, num_itrs, num_outputs, eps, use_sn=True):
        if num_svs != 1 or num_itrs != 1:
            raise NotImplementedError("ic_gan_b200 implements num_svs=1, num_itrs=1 (every IC-GAN config)")
        self.num_svs, self.num_itrs, self.eps = num_svs, num_itrs, eps
        self.register_buffer("u0", torch.randn(1, num_outputs))
        self.register_buffer("sv0", torch.ones(1))
        self._sn = ops.SNState(self, kind, use_sn)
        self._sn_table = {}

    @property
    def u(self):
        return [self.u0]

    @property
    def sv(self):
        return [self.sv0]

    def _sn_ready(self) -> "ops.SNState":
        st = self._sn
        if not st.fresh:  # stand-alone use of the layer: the owning network normally refreshes all layers at once
            with torch.no_grad():
                ops.refresh_sn([st], self.training, self.eps, self.compute_dtype, self._sn_table)
        st.fresh = False
        return st


class SNConv2d(nn.Conv2d, SN):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 num_svs=1, num_itrs=1, eps=1e-12):
        nn.Conv2d.__init__(self, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias)
        k = self.kernel_size[0]
        if self.kernel_size != (k, k) or self.stride != (1, 1) or self.padding != (k // 2, k // 2) or groups != 1 \
                or self.dilation != (1, 1) or k not in (1, 3):
            raise NotImplementedError("SNConv2d on B200: square 1x1/3x3 kernels, stride 1, 'same' padding only")
        self._sn_init("conv", num_svs, num_itrs, out_channels, eps)

    def conv_nhwc(self, x, residual=None, res_shift=0, act=ACT_NONE, out_dtype=None, stats=None, mask_input=False,
                  act_bwd_in_consumer=False):
        st = self._sn_ready()
        return ops.SNConvFn.apply(x, self.weight, self.bias, residual, st, res_shift, act,
                                  out_dtype if out_dtype is not None else x.dtype, stats, mask_input, act_bwd_in_consumer)

    def upconv_nhwc(self, x):
        """conv3x3(nearest_up2(x)) in sub-pixel form (ops.UpConvFn): x is the LOW-resolution tensor."""
        st = self._sn_ready()
        if getattr(st, "up_version", None) != ops.weight_stamp(self.weight):
            with torch.no_grad():
                st.build_up_operands()  # first use, or the weights changed since the slices were merged
        return ops.UpConvFn.apply(x, self.weight, self.bias, st)

    def downconv_nhwc(self, x, residual=None, mask_input=False):
        """avgpool2(conv3x3(x)) + residual as one stride-2 4x4 convolution (ops.DownConvFn)."""
        st = self._sn_ready()
        if getattr(st, "down_version", None) != ops.weight_stamp(self.weight):
            with torch.no_grad():
                st.build_down_operands()
        return ops.DownConvFn.apply(x, self.weight, self.bias, residual, st, mask_input)

    def bn_stats_buffer(self, x):
        """float32 [2*Cout] accumulator if this conv can emit the batch statistics of its output from its epilogue
        (tensor-core path, training, Cout % 32 == 0), else None."""
        if not ops.FUSE_BN_STATS or not self.training or x.dtype != torch.bfloat16 or self.out_channels % 32 \
                or self.in_channels % 16:
            return None
        return torch.zeros(2 * self.out_channels, device=x.device, dtype=torch.float32)

    def forward(self, x):
        xin = to_nhwc(x)
        if xin.dtype != self.compute_dtype:
            xin = xin.to(self.compute_dtype)
        return to_nchw(self.conv_nhwc(xin))


class SNLinear(nn.Linear, SN):
    def __init__(self, in_features, out_features, bias=True, num_svs=1, num_itrs=1, eps=1e-12):
        nn.Linear.__init__(self, in_features, out_features, bias)
        self._sn_init("linear", num_svs, num_itrs, out_features, eps)

    def forward(self, x):
        st = self._sn_ready()
        return ops.SNLinearFn.apply(x, self.weight, self.bias, st)


class SNEmbedding(nn.Embedding, SN):
    def __init__(self, num_embeddings, embedding_dim, padding_idx=None, max_norm=None, norm_type=2,
                 scale_grad_by_freq=False, sparse=False, _weight=None, num_svs=1, num_itrs=1, eps=1e-12):
        nn.Embedding.__init__(self, num_embeddings, embedding_dim, padding_idx, max_norm, norm_type,
                              scale_grad_by_freq, sparse, _weight)
        self._sn_init("embed", num_svs, num_itrs, num_embeddings, eps)

    def forward(self, x):
        st = self._sn_ready()
        return ops.SNEmbedFn.apply(x, self.weight, st)


# --------------------------------------------------------------------------------------------- attention
class Attention(nn.Module):
    def __init__(self, ch, which_conv=SNConv2d, name="attention"):
        super().__init__()
        self.ch = ch
        self.which_conv = which_conv
        self.theta = which_conv(ch, ch // 8, kernel_size=1, padding=0, bias=False)
        self.phi = which_conv(ch, ch // 8, kernel_size=1, padding=0, bias=False)
        self.g = which_conv(ch, ch // 2, kernel_size=1, padding=0, bias=False)
        self.o = which_conv(ch // 2, ch, kernel_size=1, padding=0, bias=False)
        self.gamma = P(torch.tensor(0.0), requires_grad=True)
        if (ch // 8) % 8:  # e.g. ch = 96: theta / phi get zero channels up to a multiple of 8 (tensor-core eligibility)
            self.theta.pad_out_to = self.phi.pad_out_to = 8

    def forward_nhwc(self, x):
        B, H, W, C = x.shape
        theta = self.theta.conv_nhwc(x)
        phi = ops.Pool2Fn.apply(self.phi.conv_nhwc(x), None, 1.0, 1)
        g = ops.Pool2Fn.apply(self.g.conv_nhwc(x), None, 1.0, 1)
        d = theta.shape[3]  # C // 8, or that rounded up to a multiple of 8 with all-zero channels (bf16 mode)
        o = ops.AttentionCoreFn.apply(theta.reshape(B, H * W, d), phi.reshape(B, H * W // 4, d),
                                      g.reshape(B, H * W // 4, C // 2))
        o = self.o.conv_nhwc(o.reshape(B, H, W, C // 2))
        return ops.ScaleAddFn.apply(o, x, self.gamma)

    def forward(self, x, y=None):
        return to_nchw(self.forward_nhwc(to_nhwc(x)))


# --------------------------------------------------------------------------------------------- normalisation
class ccbn(nn.Module):
    """Class/instance-conditional batch norm: gain = 1 + Linear(y), bias = Linear(y), F.batch_norm(momentum 0.1)."""

    def __init__(self, output_size, input_size, which_linear, eps=1e-5, momentum=0.1, cross_replica=False, mybn=False,
                 norm_style="bn"):
        super().__init__()
        if cross_replica or mybn or norm_style != "bn":
            raise NotImplementedError("ccbn on B200 implements the default norm_style='bn' path (all IC-GAN configs)")
        self.output_size, self.input_size = output_size, input_size
        self.gain = which_linear(input_size, output_size)
        self.bias = which_linear(input_size, output_size)
        self.eps, self.momentum = eps, momentum
        self.cross_replica, self.mybn, self.norm_style = cross_replica, mybn, norm_style
        self.register_buffer("stored_mean", torch.zeros(output_size))
        self.register_buffer("stored_var", torch.ones(output_size))
        self._stat_hint = {}  # last batch mean (device tensor): centres the one-pass bf16 moment kernel

    def fused(self, x_nhwc, y, relu=False, up=False, out_dtype=None, sums=None, shift=None):
        gain = 1 + self.gain(y)
        bias = self.bias(y)
        return ops.BNActFn.apply(x_nhwc, gain, bias, self.stored_mean, self.stored_var, self.training, self.eps, 0.1,
                                 relu, up, out_dtype if out_dtype is not None else x_nhwc.dtype, sums, shift,
                                 self._stat_hint)

    def forward(self, x, y):
        return to_nchw(self.fused(to_nhwc(x), y))

    def extra_repr(self):
        return f"out: {self.output_size}, in: {self.input_size}"


class bn(nn.Module):
    def __init__(self, output_size, eps=1e-5, momentum=0.1, cross_replica=False, mybn=False, **kwargs):
        super().__init__()
        if cross_replica or mybn:
            raise NotImplementedError("bn on B200 implements the default (F.batch_norm) path")
        self.output_size, self.eps, self.momentum = output_size, eps, momentum
        self.cross_replica, self.mybn = cross_replica, mybn
        self.register_buffer("stored_mean", torch.zeros(output_size))
        self.register_buffer("stored_var", torch.ones(output_size))
        self.gain = P(torch.ones(output_size), requires_grad=True)
        self.bias = P(torch.zeros(output_size), requires_grad=True)
        self._stat_hint = {}

    def fused(self, x_nhwc, relu=False, out_dtype=None, sums=None, shift=None):
        return ops.BNActFn.apply(x_nhwc, self.gain, self.bias, self.stored_mean, self.stored_var, self.training,
                                 self.eps, self.momentum, relu, False,
                                 out_dtype if out_dtype is not None else x_nhwc.dtype, sums, shift, self._stat_hint)

    def forward(self, x, y=None):
        return to_nchw(self.fused(to_nhwc(x)))


# --------------------------------------------------------------------------------------------- residual blocks
class GBlock(nn.Module):
    """bn1-ReLU-up2-conv1-bn2-ReLU-conv2 (+) up2-conv_sc. The 1x1 shortcut runs at LOW resolution and is added,
    nearest-upsampled, inside conv2's epilogue: conv1x1(up(x)) == up(conv1x1(x))."""

    def __init__(self, in_channels, out_channels, which_conv=nn.Conv2d, which_bn=bn, activation=None, upsample=None):
        super().__init__()
        if not _is_relu(activation):
            raise NotImplementedError("GBlock on B200 fuses ReLU; other activations are not implemented")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.which_conv, self.which_bn = which_conv, which_bn
        self.activation = activation
        self.upsample = upsample
        self.conv1 = which_conv(in_channels, out_channels)
        self.conv2 = which_conv(out_channels, out_channels)
        self.learnable_sc = in_channels != out_channels or upsample
        if self.learnable_sc:
            self.conv_sc = which_conv(in_channels, out_channels, kernel_size=1, padding=0)
        self.bn1 = which_bn(in_channels)
        self.bn2 = which_bn(out_channels)

    def forward_nhwc(self, x, y):
        """Batch statistics for bn2 (and for whichever batch norm consumes this block's output) are accumulated by the
        producing conv's epilogue; they travel with the tensor as the attribute `_icgan_bn = (sums, shift)`."""
        up = bool(self.upsample)
        pre = getattr(x, "_icgan_bn", (None, None))
        sub = (up and ops.SUBPIXEL_UP and x.dtype == torch.bfloat16 and self.conv1.in_channels % 16 == 0
               and self.conv1.out_channels % 16 == 0)
        if sub:
            self.conv1.sub_pixel_up = True  # refresh_sn then re-merges the slices whenever the weights change
        h = self.bn1.fused(x, y, relu=True, up=(up and not sub), sums=pre[0], shift=pre[1])
        s1 = None if sub else self.conv1.bn_stats_buffer(h)
        h = self.conv1.upconv_nhwc(h) if sub else self.conv1.conv_nhwc(h, stats=s1)
        h = self.bn2.fused(h, y, relu=True, up=False, sums=s1, shift=self.conv1.bias if s1 is not None else None)
        sc = self.conv_sc.conv_nhwc(x) if self.learnable_sc else x
        s2 = self.conv2.bn_stats_buffer(h)
        out = self.conv2.conv_nhwc(h, residual=sc, res_shift=1 if up else 0, stats=s2)
        if s2 is not None:
            out._icgan_bn = (s2, self.conv2.bias)
        return out

    def forward(self, x, y):
        return to_nchw(self.forward_nhwc(to_nhwc(x), y))


class DBlock(nn.Module):
    """[ReLU]-conv1-ReLU-conv2-[avgpool] (+) shortcut. Average pooling and the 1x1 shortcut conv commute, so the
    shortcut always pools first (4x fewer MACs); ReLU after conv1 is fused into conv1's epilogue."""

    def __init__(self, in_channels, out_channels, which_conv=SNConv2d, wide=True, preactivation=False, activation=None,
                 downsample=None):
        super().__init__()
        if not _is_relu(activation):
            raise NotImplementedError("DBlock on B200 fuses ReLU; other activations are not implemented")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.hidden_channels = out_channels if wide else in_channels
        self.which_conv = which_conv
        self.preactivation = preactivation
        self.activation = activation
        self.downsample = downsample
        self.conv1 = which_conv(in_channels, self.hidden_channels)
        self.conv2 = which_conv(self.hidden_channels, out_channels)
        self.learnable_sc = True if (in_channels != out_channels) or downsample else False
        if self.learnable_sc:
            self.conv_sc = which_conv(in_channels, out_channels, kernel_size=1, padding=0)

    def forward_nhwc(self, x):
        down = bool(self.downsample)
        # Both ReLU backwards of the block are done by the consuming conv's dgrad (gate in its epilogue): the
        # pre-activation ReLU and conv1's fused ReLU pass gradients through unchanged.
        h = ops.ReluPassFn.apply(x) if self.preactivation else x
        h = self.conv1.conv_nhwc(h, act=ACT_RELU, mask_input=self.preactivation, act_bwd_in_consumer=True)
        s = ops.Pool2Fn.apply(x, None, 0.25, 0) if down else x
        if self.learnable_sc:
            s = self.conv_sc.conv_nhwc(s)
        if down:
            if (ops.POOLED_DOWN and h.dtype == torch.bfloat16 and self.conv2.in_channels % 16 == 0
                    and self.conv2.out_channels % 8 == 0 and h.shape[1] % 2 == 0 and h.shape[2] % 2 == 0):
                self.conv2.pooled_down = True
                return self.conv2.downconv_nhwc(h, residual=s, mask_input=True)
            h = self.conv2.conv_nhwc(h, mask_input=True)
            return ops.Pool2Fn.apply(h, s, 0.25, 0)
        return self.conv2.conv_nhwc(h, residual=s, mask_input=True)

    def forward(self, x):
        xin = to_nhwc(x)
        if xin.dtype != self.conv1.compute_dtype:
            xin = xin.to(self.conv1.compute_dtype)
        return to_nchw(self.forward_nhwc(xin))
