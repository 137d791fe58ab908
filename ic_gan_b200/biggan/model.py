"""IC-GAN BigGAN ``Generator`` / ``Discriminator`` / ``G_D`` with the reference's nn.Module surface.

Mirrors ``BigGAN_PyTorch/BigGAN.py`` of facebookresearch/ic_gan: constructor keywords (Generator :88-125,
Discriminator :435-462), ``forward`` signatures (:364, :617, :655-668), attribute names read by the trainer
(``dim_z``, ``shared``, ``optim``, ``arch``, ``blocks``...) and the exact ``state_dict`` layout, so that
``trainer.py`` / ``train_fns.py`` / ``inference/utils.py`` run unchanged and reference checkpoints load strictly.
All tensor math below the [B, C] glue runs in libicgan_b200's sm_100a kernels (see layers.py / ops.py).

Extra keyword (ignored by the reference thanks to its ``**kwargs``): ``compute_dtype`` = torch.float32 ("parity mode",
exact-fp32 CUDA-core kernels) or torch.bfloat16 ("throughput mode", tcgen05 tensor cores, fp32 accumulate/statistics).
"""
from __future__ import annotations

import functools

import torch
import torch.nn as nn
import torch.nn.functional as F
import torch.optim as optim
from torch.nn import init

from . import layers
from .. import ops
from .._lib import ACT_TANH

_G_MULT = {512: ((16, 16, 8, 8, 4, 2, 1), (16, 8, 8, 4, 2, 1, 1)), 256: ((16, 16, 8, 8, 4, 2), (16, 8, 8, 4, 2, 1)),
           128: ((16, 16, 8, 4, 2), (16, 8, 4, 2, 1)), 64: ((16, 16, 8, 4), (16, 8, 4, 2)), 32: ((4, 4, 4), (4, 4, 4))}
_D_MULT = {256: ((1, 2, 4, 8, 8, 16), (1, 2, 4, 8, 8, 16, 16), (128, 64, 32, 16, 8, 4, 4), 6),
           128: ((1, 2, 4, 8, 16), (1, 2, 4, 8, 16, 16), (64, 32, 16, 8, 4, 4), 5),
           64: ((1, 2, 4, 8), (1, 2, 4, 8, 16), (32, 16, 8, 4, 4), 4),
           32: ((4, 4, 4), (4, 4, 4, 4), (16, 16, 16, 16), 2)}


def _attn_flags(spec, resolutions, lo, hi):
    wanted = {int(s) for s in str(spec).split("_") if s}
    table = {2 ** i: (2 ** i in wanted) for i in range(lo, hi)}
    return table


def G_arch(ch=64, attention="64", ksize="333333", dilation="111111"):
    """Channel/resolution tables of the generator (same dict layout as BigGAN.py:32-85)."""
    arch = {}
    for res, (mi, mo) in _G_MULT.items():
        n = len(mo)
        arch[res] = {"in_channels": [ch * m for m in mi], "out_channels": [ch * m for m in mo], "upsample": [True] * n,
                     "resolution": [8 * 2 ** i for i in range(n)],
                     "attention": _attn_flags(attention, None, 3, 3 + n)}
    return arch


def D_arch(ch=64, attention="64", ksize="333333", dilation="111111"):
    """Discriminator tables (BigGAN.py:390-432)."""
    arch = {}
    for res, (mi, mo, rr, n_down) in _D_MULT.items():
        arch[res] = {"in_channels": [3] + [ch * m for m in mi], "out_channels": [ch * m for m in mo],
                     "downsample": [i < n_down for i in range(len(mo))], "resolution": list(rr),
                     "attention": _attn_flags(attention, None, 2, 8)}
    return arch


def _init_weights(net, style):
    net.param_count = 0
    for m in net.modules():
        if isinstance(m, (nn.Conv2d, nn.Linear, nn.Embedding)):
            if style == "ortho":
                init.orthogonal_(m.weight)
            elif style == "N02":
                init.normal_(m.weight, 0, 0.02)
            elif style in ("glorot", "xavier"):
                init.xavier_uniform_(m.weight)
            else:
                print("Init style not recognized...")
            net.param_count += sum(p.data.nelement() for p in m.parameters())


class _SNNetwork(nn.Module):
    """Shared plumbing: compute dtype and the batched spectral-norm refresh that precedes every forward."""

    compute_dtype = torch.float32

    def set_compute_dtype(self, dtype):
        if dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("compute_dtype must be torch.float32 or torch.bfloat16")
        self.compute_dtype = dtype
        for m in self.modules():
            if isinstance(m, layers.SN):
                m.compute_dtype = dtype
        self._sn_cache = {}
        return self

    def _refresh_sn(self):
        if not hasattr(self, "_sn_states") or self._sn_states is None:
            self._sn_states = [m._sn for m in self.modules() if isinstance(m, layers.SN)]
            self._sn_cache = {}
        with torch.no_grad():
            ops.refresh_sn(self._sn_states, self.training, self.SN_eps, self.compute_dtype, self._sn_cache)

    def prepare_operands(self):
        """Rebuild the operand copies of every layer whose master weight changed since they were built -- what each
        forward does through refresh_sn; the graph-replay path (biggan/graphs.py) calls it because a replay runs no Python."""
        if getattr(self, "_sn_states", None):
            with torch.no_grad():
                for s in self._sn_states:
                    s.prepare()

    def _apply(self, fn, *args, **kwargs):  # .to()/.cuda() move parameters: drop cached device tables
        out = super()._apply(fn, *args, **kwargs)
        self._sn_states = None
        for m in self.modules():
            if isinstance(m, layers.SN):
                m._sn.key = None
                m._sn_table = {}
        return out


class Generator(_SNNetwork):
    def __init__(self, G_ch=64, dim_z=128, bottom_width=4, resolution=128, G_kernel_size=3, G_attn="64",
                 n_classes=1000, num_G_SVs=1, num_G_SV_itrs=1, G_shared=True, shared_dim=0, hier=False,
                 cross_replica=False, mybn=False, G_activation=nn.ReLU(inplace=False), G_lr=5e-5, G_B1=0.0,
                 G_B2=0.999, adam_eps=1e-8, BN_eps=1e-5, SN_eps=1e-12, G_mixed_precision=False, G_fp16=False,
                 G_init="ortho", skip_init=False, no_optim=False, G_param="SN", norm_style="bn", class_cond=True,
                 embedded_optimizer=True, instance_cond=False, G_shared_feat=True, shared_dim_feat=2048, **kwargs):
        super().__init__()
        if G_param != "SN" or not G_shared or (instance_cond and not G_shared_feat):
            raise NotImplementedError("ic_gan_b200 Generator: SN parameterisation with shared embeddings only")
        self.ch, self.dim_z, self.bottom_width, self.resolution = G_ch, dim_z, bottom_width, resolution
        self.kernel_size, self.attention, self.n_classes = G_kernel_size, G_attn, n_classes
        self.G_shared = G_shared
        self.shared_dim = shared_dim if shared_dim > 0 else dim_z
        self.hier, self.cross_replica, self.mybn = hier, cross_replica, mybn
        self.activation, self.init, self.G_param, self.norm_style = G_activation, G_init, G_param, norm_style
        self.BN_eps, self.SN_eps, self.fp16 = BN_eps, SN_eps, G_fp16
        self.G_shared_feat, self.shared_dim_feat = G_shared_feat, shared_dim_feat
        self.arch = G_arch(self.ch, self.attention)[resolution]

        if self.hier:  # dim_z is rounded down to a multiple of the chunk size (BigGAN.py:172-177)
            self.num_slots = len(self.arch["in_channels"]) + 1
            self.z_chunk_size = self.dim_z // self.num_slots
            self.dim_z = self.z_chunk_size * self.num_slots
        else:
            self.num_slots, self.z_chunk_size = 1, 0

        self.which_conv = functools.partial(layers.SNConv2d, kernel_size=3, padding=1, num_svs=num_G_SVs,
                                            num_itrs=num_G_SV_itrs, eps=self.SN_eps)
        self.which_linear = functools.partial(layers.SNLinear, num_svs=num_G_SVs, num_itrs=num_G_SV_itrs,
                                              eps=self.SN_eps)
        self.which_embedding = nn.Embedding  # G's class embedding is never spectrally normalised (BigGAN.py:202-204)
        bn_linear = functools.partial(self.which_linear, bias=False)
        if not class_cond and not instance_cond:
            cond_width = self.n_classes
        else:
            cond_width = self.z_chunk_size
        if class_cond:
            cond_width += self.shared_dim
        if instance_cond:
            cond_width += self.shared_dim_feat
        self.which_bn = functools.partial(layers.ccbn, which_linear=bn_linear, cross_replica=self.cross_replica,
                                          mybn=self.mybn, input_size=cond_width, norm_style=self.norm_style,
                                          eps=self.BN_eps)

        self.shared = self.which_embedding(n_classes, self.shared_dim)
        self.shared_feat = self.which_linear(2048, self.shared_dim_feat) if G_shared_feat else layers.identity()
        self.linear = self.which_linear(self.dim_z // self.num_slots,
                                        self.arch["in_channels"][0] * (self.bottom_width ** 2))

        blocks = []
        for i in range(len(self.arch["out_channels"])):
            stage = [layers.GBlock(in_channels=self.arch["in_channels"][i], out_channels=self.arch["out_channels"][i],
                                   which_conv=self.which_conv, which_bn=self.which_bn, activation=self.activation,
                                   upsample=(functools.partial(F.interpolate, scale_factor=2)
                                             if self.arch["upsample"][i] else None))]
            if self.arch["attention"][self.arch["resolution"][i]]:
                stage.append(layers.Attention(self.arch["out_channels"][i], self.which_conv))
            blocks.append(nn.ModuleList(stage))
        self.blocks = nn.ModuleList(blocks)

        self.output_layer = nn.Sequential(
            layers.bn(self.arch["out_channels"][-1], cross_replica=self.cross_replica, mybn=self.mybn),
            self.activation, self.which_conv(self.arch["out_channels"][-1], 3))

        if not skip_init:
            self.init_weights()
        self.set_compute_dtype(kwargs.get("compute_dtype", torch.float32))
        if no_optim or not embedded_optimizer:
            return
        self.lr, self.B1, self.B2, self.adam_eps = G_lr, G_B1, G_B2, adam_eps
        if G_mixed_precision:
            raise NotImplementedError("Adam16 (G_mixed_precision) is not part of the B200 hot path")
        self.optim = optim.Adam(params=self.parameters(), lr=self.lr, betas=(self.B1, self.B2), weight_decay=0,
                                eps=self.adam_eps)

    def init_weights(self):
        _init_weights(self, self.init)

    def get_condition_embeddings(self, cl=None, feat=None):
        parts = []
        if cl is not None:
            parts.append(self.shared(cl))
        if feat is not None:
            parts.append(self.shared_feat(feat))
        return torch.cat(parts, dim=-1) if parts else parts

    def forward(self, z, label=None, feats=None):
        self._refresh_sn()
        y = self.get_condition_embeddings(label, feats)
        if self.hier:
            zs = torch.split(z, self.z_chunk_size, 1)
            z = zs[0]
            ys = [torch.cat([y, zc], 1) for zc in zs[1:]]
        else:
            ys = [y] * len(self.blocks)
        h = self.linear(z)
        h = h.view(h.size(0), -1, self.bottom_width, self.bottom_width)
        h = layers.to_nhwc(h)
        if h.dtype != self.compute_dtype:
            h = h.to(self.compute_dtype)
        for i, stage in enumerate(self.blocks):
            for block in stage:
                h = block.forward_nhwc(h, ys[i]) if isinstance(block, layers.GBlock) else block.forward_nhwc(h)
        pre = getattr(h, "_icgan_bn", (None, None))
        h = self.output_layer[0].fused(h, relu=True, sums=pre[0], shift=pre[1])
        out = self.output_layer[2].conv_nhwc(h, act=ACT_TANH, out_dtype=torch.float32)
        return layers.to_nchw(out)


class Discriminator(_SNNetwork):
    def __init__(self, D_ch=64, D_wide=True, resolution=128, D_kernel_size=3, D_attn="64", n_classes=1000,
                 num_D_SVs=1, num_D_SV_itrs=1, D_activation=nn.ReLU(inplace=False), D_lr=2e-4, D_B1=0.0, D_B2=0.999,
                 adam_eps=1e-8, SN_eps=1e-12, output_dim=1, D_mixed_precision=False, D_fp16=False, D_init="ortho",
                 skip_init=False, D_param="SN", class_cond=True, embedded_optimizer=True, instance_cond=False,
                 instance_sz=2048, **kwargs):
        super().__init__()
        if D_param != "SN":
            raise NotImplementedError("ic_gan_b200 Discriminator: SN parameterisation only")
        self.ch, self.D_wide, self.resolution = D_ch, D_wide, resolution
        self.kernel_size, self.attention, self.n_classes = D_kernel_size, D_attn, n_classes
        self.activation, self.init, self.D_param = D_activation, D_init, D_param
        self.SN_eps, self.fp16 = SN_eps, D_fp16
        self.arch = D_arch(self.ch, self.attention)[resolution]
        self.which_conv = functools.partial(layers.SNConv2d, kernel_size=3, padding=1, num_svs=num_D_SVs,
                                            num_itrs=num_D_SV_itrs, eps=self.SN_eps)
        self.which_linear = functools.partial(layers.SNLinear, num_svs=num_D_SVs, num_itrs=num_D_SV_itrs,
                                              eps=self.SN_eps)
        self.which_embedding = functools.partial(layers.SNEmbedding, num_svs=num_D_SVs, num_itrs=num_D_SV_itrs,
                                                 eps=self.SN_eps)
        blocks = []
        for i in range(len(self.arch["out_channels"])):
            stage = [layers.DBlock(in_channels=self.arch["in_channels"][i], out_channels=self.arch["out_channels"][i],
                                   which_conv=self.which_conv, wide=self.D_wide, activation=self.activation,
                                   preactivation=(i > 0),
                                   downsample=(nn.AvgPool2d(2) if self.arch["downsample"][i] else None))]
            if self.arch["attention"][self.arch["resolution"][i]]:
                stage.append(layers.Attention(self.arch["out_channels"][i], self.which_conv))
            blocks.append(nn.ModuleList(stage))
        self.blocks = nn.ModuleList(blocks)
        c_last = self.arch["out_channels"][-1]
        self.linear = self.which_linear(c_last, output_dim)
        if class_cond and instance_cond:
            self.linear_feat = self.which_linear(instance_sz, c_last // 2)
            self.embed = self.which_embedding(self.n_classes, c_last // 2)
        elif class_cond:
            self.embed = self.which_embedding(self.n_classes, c_last)
        elif instance_cond:
            self.linear_feat = self.which_linear(instance_sz, c_last)
        if not skip_init:
            self.init_weights()
        self.set_compute_dtype(kwargs.get("compute_dtype", torch.float32))
        if embedded_optimizer:
            self.lr, self.B1, self.B2, self.adam_eps = D_lr, D_B1, D_B2, adam_eps
            if D_mixed_precision:
                raise NotImplementedError("Adam16 (D_mixed_precision) is not part of the B200 hot path")
            self.optim = optim.Adam(params=self.parameters(), lr=self.lr, betas=(self.B1, self.B2), weight_decay=0,
                                    eps=self.adam_eps)

    def init_weights(self):
        _init_weights(self, self.init)

    def forward(self, x, y=None, feat=None):
        self._refresh_sn()
        h = layers.to_nhwc(x)
        if h.dtype != self.compute_dtype:
            h = h.to(self.compute_dtype)
        for stage in self.blocks:
            for block in stage:
                h = block.forward_nhwc(h)
        h = ops.ReluSumPoolFn.apply(h)  # [B, C] float32, sum over H*W of ReLU (BigGAN.py:624)
        out = self.linear(h)
        proj = []
        if y is not None:
            proj.append(self.embed(y))
        if feat is not None:
            proj.append(self.linear_feat(feat))
        if proj:
            out = out + torch.sum(torch.cat(proj, dim=-1) * h, 1, keepdim=True)
        return out


class G_D(nn.Module):
    """Runs G then D on cat(fake, real) in one pass (BigGAN.py:647-711)."""

    def __init__(self, G, D, optimizer_G=None, optimizer_D=None):
        super().__init__()
        self.G, self.D = G, D
        self.optimizer_G, self.optimizer_D = optimizer_G, optimizer_D

    def forward(self, z, gy, feats_g=None, x=None, dy=None, feats=None, train_G=False, return_G_z=False,
                split_D=False, policy=False, DA=False):
        if DA:
            raise NotImplementedError("DiffAugment is off in every IC-GAN config and not part of the B200 hot path")
        with torch.set_grad_enabled(train_G):
            G_z = self.G(z, gy, feats_g)
        if split_D:
            D_fake = self.D(G_z, gy, feats_g)
            if x is not None:
                return D_fake, self.D(x, dy, feats)
            return (D_fake, G_z) if return_G_z else D_fake
        D_input = torch.cat([G_z, x], 0) if x is not None else G_z
        D_class = torch.cat([gy, dy], 0) if dy is not None else gy
        if feats_g is not None:
            D_feats = torch.cat([feats_g, feats], 0) if feats is not None else feats_g
        else:
            D_feats = None
        D_out = self.D(D_input, D_class, D_feats)
        if x is not None:
            return torch.split(D_out, [G_z.shape[0], x.shape[0]])
        return (D_out, G_z) if return_G_z else D_out
