"""StyleGAN2-ADA networks of IC-GAN on the B200 ops.

Same module surface as stylegan2_ada_pytorch/training/networks.py -- class names, constructor keywords, forward
signatures, attribute names and therefore state_dict keys / shapes (`synthesis.b64.conv0.affine.weight`,
`mapping.embed_feats.weight`, `b4.out.bias`, ...) -- so pickles' weights load with strict=True and `training_loop.py`
drives it unchanged.  The math runs on this package's ops (`bias_act`, `upfirdn2d`, `conv2d_resample`,
`modulated_conv2d`: CUDA kernels behind the C ABI, no PyTorch fallback); dense layers use torch.matmul (a plain library
GEMM, twice differentiable as the path-length regulariser needs).  `use_fp16` blocks compute in bfloat16 here
(BASELINE config 4: num_fp16_res=4 -> bf16); `conv_clamp` is kept as in the reference.

Reference lines: FullyConnectedLayer :124-160, Conv2dLayer :167-231, MappingNetwork :238-354 (IC-GAN: embed_feats
:281-282, :306-325), SynthesisLayer :361-444, ToRGBLayer :451-485, SynthesisBlock :492-635, SynthesisNetwork :642-703,
Generator :710-756, DiscriminatorBlock :763-893, MinibatchStdLayer :900-927, DiscriminatorEpilogue :934-1008,
Discriminator :1015-1101."""
from __future__ import annotations

import math

import torch
from torch import nn

from .modconv import modulated_conv2d, modulated_conv2d_act  # noqa: F401  (modulated_conv2d: reference name)
from .ops import bias_act, conv2d_resample, elementwise, upfirdn2d

LOW_PRECISION = torch.bfloat16  # what the reference's `use_fp16` blocks compute in here
_SQRT_HALF = math.sqrt(0.5)


def normalize_2nd_moment(x, dim=1, eps=1e-8):
    return x * (x.square().mean(dim=dim, keepdim=True) + eps).rsqrt()


def _filter_buffer(module, taps):
    module.register_buffer("resample_filter", upfirdn2d.setup_filter(taps))


def _cast(x, dtype, channels_last):
    """dtype cast; 4-D activations are ALWAYS channels-last on B200 (every kernel underneath is NHWC, so the reference's
    `fp16_channels_last` switch has nothing left to choose -- it is accepted and ignored)."""
    if x.ndim == 4:
        return x.to(dtype=dtype, memory_format=torch.channels_last)
    return x.to(dtype=dtype)


class FullyConnectedLayer(nn.Module):
    def __init__(self, in_features, out_features, bias=True, activation="linear", lr_multiplier=1, bias_init=0):
        super().__init__()
        self.activation = activation
        self.weight = nn.Parameter(torch.randn([out_features, in_features]) / lr_multiplier)
        self.bias = nn.Parameter(torch.full([out_features], float(bias_init))) if bias else None
        self.weight_gain = lr_multiplier / math.sqrt(in_features)
        self.bias_gain = lr_multiplier

    def forward(self, x):
        w = self.weight.to(x.dtype) * self.weight_gain
        b = self.bias
        if b is not None:
            b = b.to(x.dtype)
            if self.bias_gain != 1:
                b = b * self.bias_gain
        if self.activation == "linear" and b is not None:
            return torch.addmm(b.unsqueeze(0), x, w.t())
        return bias_act.bias_act(x.matmul(w.t()), b, act=self.activation)


class Conv2dLayer(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, bias=True, activation="linear", up=1, down=1,
                 resample_filter=[1, 3, 3, 1], conv_clamp=None, channels_last=False, trainable=True):
        super().__init__()
        self.activation, self.up, self.down, self.conv_clamp = activation, up, down, conv_clamp
        _filter_buffer(self, resample_filter)
        self.padding = kernel_size // 2
        self.weight_gain = 1 / math.sqrt(in_channels * kernel_size ** 2)
        self.act_gain = bias_act.activation_funcs[activation].def_gain
        fmt = torch.channels_last if channels_last else torch.contiguous_format
        weight = torch.randn([out_channels, in_channels, kernel_size, kernel_size]).to(memory_format=fmt)
        b = torch.zeros([out_channels]) if bias else None
        if trainable:
            self.weight = nn.Parameter(weight)
            self.bias = nn.Parameter(b) if b is not None else None
        else:  # Freeze-D layers keep their tensors as buffers (same state_dict keys)
            self.register_buffer("weight", weight)
            if b is not None:
                self.register_buffer("bias", b)
            else:
                self.bias = None

    def forward(self, x, gain=1):
        plain = self.activation == "linear" and self.bias is None and self.conv_clamp is None
        # a linear layer without bias or clamp (the residual skip path, gain sqrt(1/2)): the output gain goes into the
        # weights instead of costing a pass over the activation
        w = (self.weight * (self.weight_gain * self.act_gain * gain if plain else self.weight_gain)).to(x.dtype)
        b = self.bias.to(x.dtype) if self.bias is not None else None
        x = conv2d_resample.conv2d_resample(x=x, w=w, f=self.resample_filter, up=self.up, down=self.down,
                                            padding=self.padding, flip_weight=(self.up == 1))
        if plain:
            return x
        clamp = self.conv_clamp * gain if self.conv_clamp is not None else None
        # bias + activation + gain + clamp in one pass, and ONE pass again for (dx, dbias) in the first-order backward
        return elementwise.mod_bias_act(x, bias=b, act=self.activation, gain=self.act_gain * gain, clamp=clamp)


class MappingNetwork(nn.Module):
    def __init__(self, z_dim, c_dim, h_dim, w_dim, num_ws, num_layers=8, embed_features=None, embed_features_feat=None,
                 layer_features=None, activation="lrelu", lr_multiplier=0.01, w_avg_beta=0.995):
        super().__init__()
        self.z_dim, self.c_dim, self.h_dim, self.w_dim = z_dim, c_dim, h_dim, w_dim
        self.num_ws, self.num_layers, self.w_avg_beta = num_ws, num_layers, w_avg_beta
        e_c = 0 if c_dim == 0 else (w_dim if embed_features is None else embed_features)
        e_h = 0 if h_dim == 0 else (w_dim if embed_features_feat is None else embed_features_feat)
        hidden = w_dim if layer_features is None else layer_features
        widths = [z_dim + e_c + e_h] + [hidden] * (num_layers - 1) + [w_dim]
        if c_dim > 0:
            self.embed = FullyConnectedLayer(c_dim, e_c)
        if h_dim > 0:
            self.embed_feats = FullyConnectedLayer(h_dim, e_h)
        for i in range(num_layers):
            setattr(self, f"fc{i}", FullyConnectedLayer(widths[i], widths[i + 1], activation=activation,
                                                        lr_multiplier=lr_multiplier))
        if num_ws is not None and w_avg_beta is not None:
            self.register_buffer("w_avg", torch.zeros([w_dim]))

    def forward(self, z, c, h, truncation_psi=1, truncation_cutoff=None, skip_w_avg_update=False):
        parts = []
        if self.z_dim > 0:
            assert z.shape[1] == self.z_dim
            parts.append(normalize_2nd_moment(z.to(torch.float32)))
        cond = []
        if self.c_dim > 0:
            assert c.shape[1] == self.c_dim
            cond.append(self.embed(c.to(torch.float32)))
        if self.h_dim > 0:
            assert h.shape[1] == self.h_dim
            cond.append(self.embed_feats(h.to(torch.float32)))
        if cond:  # label and instance embeddings are normalised jointly (networks.py:306-317)
            parts.append(normalize_2nd_moment(torch.cat(cond, dim=1) if len(cond) > 1 else cond[0]))
        x = torch.cat(parts, dim=1) if len(parts) > 1 else parts[0]
        for i in range(self.num_layers):
            x = getattr(self, f"fc{i}")(x)
        if self.w_avg_beta is not None and self.training and not skip_w_avg_update:
            self.w_avg.copy_(x.detach().mean(dim=0).lerp(self.w_avg, self.w_avg_beta))
        if self.num_ws is not None:
            x = x.unsqueeze(1).repeat([1, self.num_ws, 1])
        if truncation_psi != 1:
            assert self.w_avg_beta is not None
            if self.num_ws is None or truncation_cutoff is None:
                x = self.w_avg.lerp(x, truncation_psi)
            else:
                x[:, :truncation_cutoff] = self.w_avg.lerp(x[:, :truncation_cutoff], truncation_psi)
        return x


class SynthesisLayer(nn.Module):
    def __init__(self, in_channels, out_channels, w_dim, resolution, kernel_size=3, up=1, use_noise=True,
                 activation="lrelu", resample_filter=[1, 3, 3, 1], conv_clamp=None, channels_last=False):
        super().__init__()
        self.resolution, self.up, self.use_noise = resolution, up, use_noise
        self.activation, self.conv_clamp = activation, conv_clamp
        _filter_buffer(self, resample_filter)
        self.padding = kernel_size // 2
        self.act_gain = bias_act.activation_funcs[activation].def_gain
        self.affine = FullyConnectedLayer(w_dim, in_channels, bias_init=1)
        fmt = torch.channels_last if channels_last else torch.contiguous_format
        self.weight = nn.Parameter(torch.randn([out_channels, in_channels, kernel_size, kernel_size]).to(memory_format=fmt))
        if use_noise:
            self.register_buffer("noise_const", torch.randn([resolution, resolution]))
            self.noise_strength = nn.Parameter(torch.zeros([]))
        self.bias = nn.Parameter(torch.zeros([out_channels]))

    def forward(self, x, w, noise_mode="random", fused_modconv=True, gain=1):
        assert noise_mode in ["random", "const", "none"]
        assert x.shape[1] == self.weight.shape[1] and x.shape[2] == x.shape[3] == self.resolution // self.up
        styles = self.affine(w)
        noise = None
        if self.use_noise and noise_mode == "random":
            noise = torch.randn([x.shape[0], 1, self.resolution, self.resolution], device=x.device) * self.noise_strength
        elif self.use_noise and noise_mode == "const":
            noise = self.noise_const * self.noise_strength
        clamp = self.conv_clamp * gain if self.conv_clamp is not None else None
        # x*styles -> conv -> one pass of (*dcoefs + noise + bias -> act -> gain -> clamp)   (networks.py:77-95, :441-444)
        return modulated_conv2d_act(x=x, weight=self.weight, styles=styles, noise=noise, bias=self.bias,
                                    act=self.activation, gain=self.act_gain * gain, clamp=clamp, up=self.up,
                                    padding=self.padding, resample_filter=self.resample_filter,
                                    flip_weight=(self.up == 1))


class ToRGBLayer(nn.Module):
    def __init__(self, in_channels, out_channels, w_dim, kernel_size=1, conv_clamp=None, channels_last=False):
        super().__init__()
        self.conv_clamp = conv_clamp
        self.affine = FullyConnectedLayer(w_dim, in_channels, bias_init=1)
        fmt = torch.channels_last if channels_last else torch.contiguous_format
        self.weight = nn.Parameter(torch.randn([out_channels, in_channels, kernel_size, kernel_size]).to(memory_format=fmt))
        self.bias = nn.Parameter(torch.zeros([out_channels]))
        self.weight_gain = 1 / math.sqrt(in_channels * kernel_size ** 2)

    def forward(self, x, w, fused_modconv=True):
        styles = self.affine(w) * self.weight_gain
        return modulated_conv2d_act(x=x, weight=self.weight, styles=styles, bias=self.bias, clamp=self.conv_clamp,
                                    demodulate=False)


class SynthesisBlock(nn.Module):
    def __init__(self, in_channels, out_channels, w_dim, resolution, img_channels, is_last, architecture="skip",
                 resample_filter=[1, 3, 3, 1], conv_clamp=None, use_fp16=False, fp16_channels_last=False, **layer_kwargs):
        assert architecture in ["orig", "skip", "resnet"]
        super().__init__()
        self.in_channels, self.w_dim, self.resolution, self.img_channels = in_channels, w_dim, resolution, img_channels
        self.is_last, self.architecture, self.use_fp16 = is_last, architecture, use_fp16
        self.channels_last = use_fp16 and fp16_channels_last
        _filter_buffer(self, resample_filter)
        self.num_conv = self.num_torgb = 0
        common = dict(w_dim=w_dim, resolution=resolution, conv_clamp=conv_clamp, channels_last=self.channels_last,
                      **layer_kwargs)
        if in_channels == 0:
            self.const = nn.Parameter(torch.randn([out_channels, resolution, resolution]))
        else:
            self.conv0 = SynthesisLayer(in_channels, out_channels, up=2, resample_filter=resample_filter, **common)
            self.num_conv += 1
        self.conv1 = SynthesisLayer(out_channels, out_channels, **common)
        self.num_conv += 1
        if is_last or architecture == "skip":
            self.torgb = ToRGBLayer(out_channels, img_channels, w_dim=w_dim, conv_clamp=conv_clamp,
                                    channels_last=self.channels_last)
            self.num_torgb += 1
        if in_channels != 0 and architecture == "resnet":
            self.skip = Conv2dLayer(in_channels, out_channels, kernel_size=1, bias=False, up=2,
                                    resample_filter=resample_filter, channels_last=self.channels_last)

    def forward(self, x, img, ws, force_fp32=False, fused_modconv=None, **layer_kwargs):
        assert ws.shape[1] == self.num_conv + self.num_torgb and ws.shape[2] == self.w_dim
        latents = iter(ws.unbind(dim=1))
        low = self.use_fp16 and not force_fp32
        dtype = LOW_PRECISION if low else torch.float32
        cl = self.channels_last and not force_fp32
        if fused_modconv is None:  # training always takes the non-fused form (networks.py:589-594)
            fused_modconv = (not self.training) and (dtype == torch.float32 or int(x.shape[0]) == 1)
        if self.in_channels == 0:
            x = _cast(self.const, dtype, cl).unsqueeze(0).repeat([ws.shape[0], 1, 1, 1])
            x = self.conv1(x, next(latents), fused_modconv=fused_modconv, **layer_kwargs)
        else:
            assert x.shape[1] == self.in_channels and x.shape[2] == self.resolution // 2
            x = _cast(x, dtype, cl)
            if self.architecture == "resnet":
                y = self.skip(x, gain=_SQRT_HALF)
                x = self.conv0(x, next(latents), fused_modconv=fused_modconv, **layer_kwargs)
                x = self.conv1(x, next(latents), fused_modconv=fused_modconv, gain=_SQRT_HALF, **layer_kwargs)
                x = y + x  # (the skip output is a view produced by a custom Function: no in-place update)
            else:
                x = self.conv0(x, next(latents), fused_modconv=fused_modconv, **layer_kwargs)
                x = self.conv1(x, next(latents), fused_modconv=fused_modconv, **layer_kwargs)
        if img is not None:
            img = upfirdn2d.upsample2d(img, self.resample_filter)
        if self.is_last or self.architecture == "skip":
            y = self.torgb(x, next(latents), fused_modconv=fused_modconv)
            y = y.to(dtype=torch.float32, memory_format=torch.contiguous_format)
            img = img.add_(y) if img is not None else y
        assert x.dtype == dtype and (img is None or img.dtype == torch.float32)
        return x, img


class SynthesisNetwork(nn.Module):
    def __init__(self, w_dim, img_resolution, img_channels, channel_base=32768, channel_max=512, num_fp16_res=0,
                 **block_kwargs):
        assert img_resolution >= 4 and img_resolution & (img_resolution - 1) == 0
        super().__init__()
        self.w_dim, self.img_resolution, self.img_channels = w_dim, img_resolution, img_channels
        self.img_resolution_log2 = int(math.log2(img_resolution))
        self.block_resolutions = [2 ** i for i in range(2, self.img_resolution_log2 + 1)]
        width = {res: min(channel_base // res, channel_max) for res in self.block_resolutions}
        low_from = max(2 ** (self.img_resolution_log2 + 1 - num_fp16_res), 8)
        self.num_ws = 0
        for res in self.block_resolutions:
            block = SynthesisBlock(width[res // 2] if res > 4 else 0, width[res], w_dim=w_dim, resolution=res,
                                   img_channels=img_channels, is_last=(res == img_resolution),
                                   use_fp16=(res >= low_from), **block_kwargs)
            self.num_ws += block.num_conv + (block.num_torgb if res == img_resolution else 0)
            setattr(self, f"b{res}", block)

    def forward(self, ws, **block_kwargs):
        assert ws.shape[1] == self.num_ws and ws.shape[2] == self.w_dim
        ws = ws.to(torch.float32)
        x = img = None
        first = 0
        for res in self.block_resolutions:  # a block's toRGB shares its latent with the next block's first conv
            block = getattr(self, f"b{res}")
            x, img = block(x, img, ws.narrow(1, first, block.num_conv + block.num_torgb), **block_kwargs)
            first += block.num_conv
        return img


class Generator(nn.Module):
    def __init__(self, z_dim, c_dim, h_dim, w_dim, img_resolution, img_channels, mapping_kwargs={}, synthesis_kwargs={}):
        super().__init__()
        self.z_dim, self.c_dim, self.h_dim, self.w_dim = z_dim, c_dim, h_dim, w_dim
        self.img_resolution, self.img_channels = img_resolution, img_channels
        self.synthesis = SynthesisNetwork(w_dim=w_dim, img_resolution=img_resolution, img_channels=img_channels,
                                          **synthesis_kwargs)
        self.num_ws = self.synthesis.num_ws
        self.mapping = MappingNetwork(z_dim=z_dim, c_dim=c_dim, h_dim=h_dim, w_dim=w_dim, num_ws=self.num_ws,
                                      **mapping_kwargs)

    def forward(self, z, c, feats, truncation_psi=1, truncation_cutoff=None, **synthesis_kwargs):
        ws = self.mapping(z, c, feats, truncation_psi=truncation_psi, truncation_cutoff=truncation_cutoff)
        return self.synthesis(ws, **synthesis_kwargs)


class DiscriminatorBlock(nn.Module):
    def __init__(self, in_channels, tmp_channels, out_channels, resolution, img_channels, first_layer_idx,
                 architecture="resnet", activation="lrelu", resample_filter=[1, 3, 3, 1], conv_clamp=None, use_fp16=False,
                 fp16_channels_last=False, freeze_layers=0):
        assert in_channels in [0, tmp_channels] and architecture in ["orig", "skip", "resnet"]
        super().__init__()
        self.in_channels, self.resolution, self.img_channels = in_channels, resolution, img_channels
        self.first_layer_idx, self.architecture, self.use_fp16 = first_layer_idx, architecture, use_fp16
        self.channels_last = use_fp16 and fp16_channels_last
        _filter_buffer(self, resample_filter)
        self.num_layers = 0

        def next_trainable():  # Freeze-D: layers are numbered in construction order
            idx = self.first_layer_idx + self.num_layers
            self.num_layers += 1
            return idx >= freeze_layers

        cl = self.channels_last
        if in_channels == 0 or architecture == "skip":
            self.fromrgb = Conv2dLayer(img_channels, tmp_channels, kernel_size=1, activation=activation,
                                       trainable=next_trainable(), conv_clamp=conv_clamp, channels_last=cl)
        self.conv0 = Conv2dLayer(tmp_channels, tmp_channels, kernel_size=3, activation=activation,
                                 trainable=next_trainable(), conv_clamp=conv_clamp, channels_last=cl)
        self.conv1 = Conv2dLayer(tmp_channels, out_channels, kernel_size=3, activation=activation, down=2,
                                 trainable=next_trainable(), resample_filter=resample_filter, conv_clamp=conv_clamp,
                                 channels_last=cl)
        if architecture == "resnet":
            self.skip = Conv2dLayer(tmp_channels, out_channels, kernel_size=1, bias=False, down=2,
                                    trainable=next_trainable(), resample_filter=resample_filter, channels_last=cl)

    def forward(self, x, img, force_fp32=False):
        low = self.use_fp16 and not force_fp32
        dtype = LOW_PRECISION if low else torch.float32
        cl = self.channels_last and not force_fp32
        if x is not None:
            assert x.shape[1] == self.in_channels and x.shape[2] == self.resolution
            x = _cast(x, dtype, cl)
        if self.in_channels == 0 or self.architecture == "skip":
            assert img.shape[1] == self.img_channels and img.shape[2] == self.resolution
            img = _cast(img, dtype, cl)
            y = self.fromrgb(img)
            x = x + y if x is not None else y
            img = upfirdn2d.downsample2d(img, self.resample_filter) if self.architecture == "skip" else None
        if self.architecture == "resnet":
            y = self.skip(x, gain=_SQRT_HALF)
            x = self.conv1(self.conv0(x), gain=_SQRT_HALF)
            x = y + x  # (the skip output is a view produced by a custom Function: no in-place update)
        else:
            x = self.conv1(self.conv0(x))
        assert x.dtype == dtype
        return x, img


class MinibatchStdLayer(nn.Module):
    def __init__(self, group_size, num_channels=1):
        super().__init__()
        self.group_size, self.num_channels = group_size, num_channels

    def forward(self, x):
        N, C, H, W = x.shape
        G = min(int(self.group_size), N) if self.group_size is not None else N
        F = self.num_channels
        y = x.reshape(G, -1, F, C // F, H, W)          # [G, n, F, c, H, W]: n groups of G samples
        y = y - y.mean(dim=0)
        y = (y.square().mean(dim=0) + 1e-8).sqrt()     # stddev over the group
        y = y.mean(dim=[2, 3, 4]).reshape(-1, F, 1, 1)  # [n, F, 1, 1]
        return torch.cat([x, y.repeat(G, 1, H, W)], dim=1)


class DiscriminatorEpilogue(nn.Module):
    def __init__(self, in_channels, cmap_dim, resolution, img_channels, architecture="resnet", mbstd_group_size=4,
                 mbstd_num_channels=1, activation="lrelu", conv_clamp=None):
        assert architecture in ["orig", "skip", "resnet"]
        super().__init__()
        self.in_channels, self.cmap_dim, self.resolution = in_channels, cmap_dim, resolution
        self.img_channels, self.architecture = img_channels, architecture
        if architecture == "skip":
            self.fromrgb = Conv2dLayer(img_channels, in_channels, kernel_size=1, activation=activation)
        self.mbstd = MinibatchStdLayer(mbstd_group_size, mbstd_num_channels) if mbstd_num_channels > 0 else None
        self.conv = Conv2dLayer(in_channels + mbstd_num_channels, in_channels, kernel_size=3, activation=activation,
                                conv_clamp=conv_clamp)
        self.fc = FullyConnectedLayer(in_channels * resolution ** 2, in_channels, activation=activation)
        self.out = FullyConnectedLayer(in_channels, 1 if cmap_dim == 0 else cmap_dim)

    def forward(self, x, img, cmap, force_fp32=False):
        assert x.shape[1] == self.in_channels and x.shape[2] == self.resolution
        x = _cast(x, torch.float32, False)
        if self.architecture == "skip":
            x = x + self.fromrgb(_cast(img, torch.float32, False))
        if self.mbstd is not None:
            x = self.mbstd(x)
        x = self.out(self.fc(self.conv(x).flatten(1)))
        if self.cmap_dim > 0:  # projection on the mapped conditioning (networks.py:1003-1005)
            assert cmap.shape[1] == self.cmap_dim
            x = (x * cmap).sum(dim=1, keepdim=True) * (1 / math.sqrt(self.cmap_dim))
        return x


class Discriminator(nn.Module):
    def __init__(self, c_dim, h_dim, img_resolution, img_channels, architecture="resnet", channel_base=32768,
                 channel_max=512, num_fp16_res=0, conv_clamp=None, cmap_dim=None, block_kwargs={}, mapping_kwargs={},
                 epilogue_kwargs={}):
        super().__init__()
        self.c_dim, self.h_dim, self.img_resolution, self.img_channels = c_dim, h_dim, img_resolution, img_channels
        self.img_resolution_log2 = int(math.log2(img_resolution))
        self.block_resolutions = [2 ** i for i in range(self.img_resolution_log2, 2, -1)]
        width = {res: min(channel_base // res, channel_max) for res in self.block_resolutions + [4]}
        low_from = max(2 ** (self.img_resolution_log2 + 1 - num_fp16_res), 8)
        if cmap_dim is None:
            cmap_dim = width[4]
        if c_dim == 0 and h_dim == 0:
            cmap_dim = 0
        common = dict(img_channels=img_channels, architecture=architecture, conv_clamp=conv_clamp)
        layer_idx = 0
        for res in self.block_resolutions:
            block = DiscriminatorBlock(width[res] if res < img_resolution else 0, width[res], width[res // 2],
                                       resolution=res, first_layer_idx=layer_idx, use_fp16=(res >= low_from),
                                       **block_kwargs, **common)
            setattr(self, f"b{res}", block)
            layer_idx += block.num_layers
        if c_dim > 0 or h_dim > 0:
            self.mapping = MappingNetwork(z_dim=0, c_dim=c_dim, h_dim=h_dim, w_dim=cmap_dim, num_ws=None, w_avg_beta=None,
                                          **mapping_kwargs)
        self.b4 = DiscriminatorEpilogue(width[4], cmap_dim=cmap_dim, resolution=4, **epilogue_kwargs, **common)

    def forward(self, img, c, h, **block_kwargs):
        x = None
        for res in self.block_resolutions:
            x, img = getattr(self, f"b{res}")(x, img, **block_kwargs)
        cmap = self.mapping(None, c, h) if (self.c_dim > 0 or self.h_dim > 0) else None
        return self.b4(x, img, cmap)
