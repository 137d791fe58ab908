"""CPU oracle for the StyleGAN2-ADA (IC-GAN variant) networks -- TEST INFRASTRUCTURE ONLY (never imported by
ic_gan_b200/).  SURVEY.md section 8 rows a21 (mapping / synthesis / generator) and a22 (discriminator): the oracle the
next round's B200 modules are built against.

Functional float32 restatement of stylegan2_ada_pytorch/training/networks.py on REFERENCE-LAYOUT state_dicts (same
key names and shapes as `Generator.state_dict()` / `Discriminator.state_dict()`), composed from the op oracles in
oracle/stylegan_ops_oracle.py.  Every function cites the reference lines it follows.  Pinned against the live
reference by oracle/make_golden_stylegan_nets.py (images, logits, w_avg update, and parameter gradients of the
non-saturating logistic losses of training/loss.py:96-100,126-150)."""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np
import torch
from torch import Tensor

from oracle import stylegan_ops_oracle as ops

SD = Dict[str, Tensor]


@dataclass
class StyleGANConfig:
    """Constructor arguments of networks.Generator (:710-756) / networks.Discriminator (:1015-1101) that shape the
    computation (IC-GAN: c_dim 0 or label dim, h_dim = instance feature dim)."""
    z_dim: int = 512
    c_dim: int = 0
    h_dim: int = 2048
    w_dim: int = 512
    img_resolution: int = 256
    img_channels: int = 3
    channel_base: int = 16384
    channel_max: int = 512
    map_layers: int = 2          # mapping_kwargs.num_layers of G (train.py cfg 'auto': map=2)
    d_map_layers: int = 8        # MappingNetwork default inside D (:258) unless mapping_kwargs overrides it
    conv_clamp: Optional[float] = 256.0
    mbstd_group_size: Optional[int] = 4
    mbstd_num_channels: int = 1
    w_avg_beta: float = 0.995
    lr_multiplier: float = 0.01  # mapping fc layers (:251)
    resample_filter: tuple = (1, 3, 3, 1)
    block_resolutions: list = field(init=False)

    def __post_init__(self):
        lg = int(math.log2(self.img_resolution))
        assert 2 ** lg == self.img_resolution and self.img_resolution >= 4
        self.block_resolutions = [2 ** i for i in range(2, lg + 1)]

    def channels(self, res: int) -> int:  # :662-664, :1046-1049
        return min(self.channel_base // res, self.channel_max)

    @property
    def num_ws(self) -> int:  # :668-687: one per conv + the last block's torgb
        n = 0
        for res in self.block_resolutions:
            n += 1 if res == 4 else 2
        return n + 1


def _filter(cfg: StyleGANConfig) -> Tensor:
    return ops.setup_filter(list(cfg.resample_filter))


def normalize_2nd_moment(x: Tensor, dim: int = 1, eps: float = 1e-8) -> Tensor:  # networks.py:29-30
    return x * (x.square().mean(dim=dim, keepdim=True) + eps).rsqrt()


def fully_connected(x: Tensor, sd: SD, p: str, activation: str = "linear", lr_multiplier: float = 1.0) -> Tensor:
    """FullyConnectedLayer.forward :147-160: weight * lr_mul/sqrt(in), bias * lr_mul, bias_act."""
    w = sd[p + ".weight"]
    w = w * (lr_multiplier / math.sqrt(w.shape[1]))
    b = sd.get(p + ".bias")
    if b is not None and lr_multiplier != 1:
        b = b * lr_multiplier
    return ops.bias_act(x.matmul(w.t()), b, act=activation)


def mapping(sd: SD, p: str, cfg: StyleGANConfig, z: Optional[Tensor], c: Optional[Tensor], h: Optional[Tensor], *,
            z_dim: int, num_layers: int, num_ws: Optional[int], truncation_psi: float = 1.0,
            truncation_cutoff: Optional[int] = None, training: bool = False, w_avg_beta: Optional[float] = None,
            buffers_out: Optional[dict] = None) -> Tensor:
    """MappingNetwork.forward :296-354 (IC-GAN adds embed_feats :281-282 and the concatenation order embed(c) |
    embed_feats(h) :306-317)."""
    x = None
    if z_dim > 0:
        x = normalize_2nd_moment(z.float())
    if cfg.c_dim > 0 and cfg.h_dim > 0:
        y = torch.cat([fully_connected(c.float(), sd, p + ".embed"), fully_connected(h.float(), sd, p + ".embed_feats")], 1)
        y = normalize_2nd_moment(y)
        x = torch.cat([x, y], 1) if x is not None else y
    elif cfg.c_dim > 0:
        y = normalize_2nd_moment(fully_connected(c.float(), sd, p + ".embed"))
        x = torch.cat([x, y], 1) if x is not None else y
    elif cfg.h_dim > 0:
        y = normalize_2nd_moment(fully_connected(h.float(), sd, p + ".embed_feats"))
        x = torch.cat([x, y], 1) if x is not None else y
    for i in range(num_layers):
        x = fully_connected(x, sd, f"{p}.fc{i}", activation="lrelu", lr_multiplier=cfg.lr_multiplier)
    if w_avg_beta is not None and training and buffers_out is not None:  # :330-335
        buffers_out[p + ".w_avg"] = x.detach().mean(dim=0).lerp(sd[p + ".w_avg"], w_avg_beta)
    if num_ws is not None:
        x = x.unsqueeze(1).repeat([1, num_ws, 1])
    if truncation_psi != 1:  # :343-353
        w_avg = sd[p + ".w_avg"]
        if num_ws is None or truncation_cutoff is None:
            x = w_avg.lerp(x, truncation_psi)
        else:
            x = torch.cat([w_avg.lerp(x[:, :truncation_cutoff], truncation_psi), x[:, truncation_cutoff:]], 1)
    return x


def synthesis_layer(x: Tensor, w: Tensor, sd: SD, p: str, cfg: StyleGANConfig, resolution: int, up: int,
                    noise_mode: str, gain: float = 1.0) -> Tensor:
    """SynthesisLayer.forward :405-444: affine (bias_init 1) -> modulated conv (+noise) -> bias_act lrelu, gain, clamp."""
    styles = fully_connected(w, sd, p + ".affine")
    noise = None
    if noise_mode == "random":  # same torch.randn call (global stream) as :413-419
        noise = torch.randn([x.shape[0], 1, resolution, resolution]) * sd[p + ".noise_strength"]
    elif noise_mode == "const":
        noise = sd[p + ".noise_const"] * sd[p + ".noise_strength"]
    weight = sd[p + ".weight"]
    y = _modconv(x, weight, styles, noise, up, weight.shape[-1] // 2, _filter(cfg), flip_weight=(up == 1))
    act_gain = math.sqrt(2) * gain
    clamp = cfg.conv_clamp * gain if cfg.conv_clamp is not None else None
    return ops.bias_act(y, sd[p + ".bias"], act="lrelu", gain=act_gain, clamp=clamp)


def _modconv(x, weight, styles, noise, up, padding, f, flip_weight=True, demodulate=True):
    """modulated_conv2d :37-117 per sample (mathematical form; ops oracle) with the flip_weight convention of
    conv2d_resample (:79-216: flip_weight=False = true convolution, used by the up=2 layers)."""
    outs = []
    for n in range(x.shape[0]):
        wn = weight * styles[n].reshape(1, -1, 1, 1)
        if demodulate:
            wn = wn * (wn.square().sum(dim=[1, 2, 3], keepdim=True) + 1e-8).rsqrt()
        outs.append(ops.conv2d_resample(x[n:n + 1], wn, f, up=up, padding=padding, flip_weight=flip_weight))
    y = torch.cat(outs, 0)
    return y if noise is None else y + noise


def to_rgb(x: Tensor, w: Tensor, sd: SD, p: str, cfg: StyleGANConfig) -> Tensor:
    """ToRGBLayer.forward :474-485: styles * 1/sqrt(Cin), no demodulation, linear bias_act with clamp."""
    weight = sd[p + ".weight"]
    styles = fully_connected(w, sd, p + ".affine") * (1.0 / math.sqrt(weight.shape[1] * weight.shape[2] ** 2))
    y = _modconv(x, weight, styles, None, 1, 0, None, demodulate=False)
    return ops.bias_act(y, sd[p + ".bias"], clamp=cfg.conv_clamp)


def upsample2d(x: Tensor, f: Tensor, up: int = 2) -> Tensor:  # upfirdn2d.py:392-436
    fw = fh = int(f.shape[-1])
    pad = [(fw + up - 1) // 2, (fw - up) // 2, (fh + up - 1) // 2, (fh - up) // 2]
    return ops.upfirdn2d(x, f, up=up, padding=pad, gain=up * up)


def synthesis(sd: SD, p: str, cfg: StyleGANConfig, ws: Tensor, noise_mode: str = "const") -> Tensor:
    """SynthesisNetwork.forward :689-703 + SynthesisBlock.forward :575-635 ('skip' architecture: RGB accumulated)."""
    ws = ws.float()
    x = img = None
    w_idx = 0
    for res in cfg.block_resolutions:
        b = f"{p}.b{res}"
        num_conv = 1 if res == 4 else 2
        cur = ws.narrow(1, w_idx, num_conv + 1)  # convs + torgb; the torgb latent is shared with the next block
        w_idx += num_conv
        it = iter(cur.unbind(dim=1))
        if res == 4:
            x = sd[b + ".const"].unsqueeze(0).repeat([ws.shape[0], 1, 1, 1])
            x = synthesis_layer(x, next(it), sd, b + ".conv1", cfg, res, 1, noise_mode)
        else:
            x = synthesis_layer(x, next(it), sd, b + ".conv0", cfg, res, 2, noise_mode)
            x = synthesis_layer(x, next(it), sd, b + ".conv1", cfg, res, 1, noise_mode)
        if img is not None:
            img = upsample2d(img, _filter(cfg))
        y = to_rgb(x, next(it), sd, b + ".torgb", cfg)
        img = img + y if img is not None else y
    return img


def generator(sd: SD, cfg: StyleGANConfig, z: Tensor, c: Optional[Tensor], feats: Optional[Tensor], *,
              truncation_psi: float = 1.0, truncation_cutoff: Optional[int] = None, noise_mode: str = "const",
              training: bool = False, buffers_out: Optional[dict] = None) -> Tensor:
    """Generator.forward :744-756."""
    ws = mapping(sd, "mapping", cfg, z, c, feats, z_dim=cfg.z_dim, num_layers=cfg.map_layers, num_ws=cfg.num_ws,
                 truncation_psi=truncation_psi, truncation_cutoff=truncation_cutoff, training=training,
                 w_avg_beta=cfg.w_avg_beta, buffers_out=buffers_out)
    return synthesis(sd, "synthesis", cfg, ws, noise_mode)


# ------------------------------------------------------------------------------------------------- discriminator
def conv2d_layer(x: Tensor, sd: SD, p: str, cfg: StyleGANConfig, activation: str = "linear", up: int = 1, down: int = 1,
                 gain: float = 1.0, clamp: bool = True) -> Tensor:
    """Conv2dLayer.forward :214-231."""
    w = sd[p + ".weight"]
    k = w.shape[-1]
    w = w * (1.0 / math.sqrt(w.shape[1] * k * k))
    y = ops.conv2d_resample(x, w, _filter(cfg), up=up, down=down, padding=k // 2, flip_weight=(up == 1))
    act_gain = ops.ACTS[activation][2] * gain
    cl = cfg.conv_clamp * gain if (clamp and cfg.conv_clamp is not None) else None
    return ops.bias_act(y, sd.get(p + ".bias"), act=activation, gain=act_gain, clamp=cl)


def minibatch_std(x: Tensor, group_size: Optional[int], num_channels: int = 1) -> Tensor:
    """MinibatchStdLayer.forward :906-927."""
    N, C, H, W = x.shape
    G = min(group_size, N) if group_size is not None else N
    Fc = num_channels
    c = C // Fc
    y = x.reshape(G, -1, Fc, c, H, W)
    y = y - y.mean(dim=0)
    y = y.square().mean(dim=0)
    y = (y + 1e-8).sqrt()
    y = y.mean(dim=[2, 3, 4])
    y = y.reshape(-1, Fc, 1, 1).repeat(G, 1, H, W)
    return torch.cat([x, y], dim=1)


def discriminator(sd: SD, cfg: StyleGANConfig, img: Tensor, c: Optional[Tensor], h: Optional[Tensor]) -> Tensor:
    """Discriminator.forward :1091-1101 with resnet blocks (:848-893) and the epilogue (:978-1008); the conditioning
    enters through mapping(None, c, h) -> cmap and the projection :1003-1005."""
    x = None
    s = math.sqrt(0.5)
    for res in cfg.block_resolutions[:0:-1]:  # img_resolution ... 8
        b = f"b{res}"
        if res == cfg.img_resolution:
            x = conv2d_layer(img.float(), sd, b + ".fromrgb", cfg, activation="lrelu")
        y = conv2d_layer(x, sd, b + ".skip", cfg, down=2, gain=s, clamp=False)  # skip has no conv_clamp (:836-846)
        x = conv2d_layer(x, sd, b + ".conv0", cfg, activation="lrelu")
        x = conv2d_layer(x, sd, b + ".conv1", cfg, activation="lrelu", down=2, gain=s)
        x = y + x
    cmap = None
    cmap_dim = 0
    if cfg.c_dim > 0 or cfg.h_dim > 0:
        cmap_dim = cfg.channels(4)
        cmap = mapping(sd, "mapping", cfg, None, c, h, z_dim=0, num_layers=cfg.d_map_layers, num_ws=None)
    if cfg.mbstd_num_channels > 0:
        x = minibatch_std(x, cfg.mbstd_group_size, cfg.mbstd_num_channels)
    x = conv2d_layer(x, sd, "b4.conv", cfg, activation="lrelu")
    x = fully_connected(x.flatten(1), sd, "b4.fc", activation="lrelu")
    x = fully_connected(x, sd, "b4.out")
    if cmap_dim > 0:
        x = (x * cmap).sum(dim=1, keepdim=True) * (1.0 / math.sqrt(cmap_dim))
    return x


# ------------------------------------------------------------------------------------------------- synthetic weights
def synth_state_dict(shapes: Dict[str, tuple], seed: int) -> SD:
    """Deterministic non-trivial parameters (numpy PCG64 keyed by sorted key order): unit-variance weights, small biases,
    non-zero noise strengths -- the reference's own init leaves noise_strength = 0 and biases = 0, which would hide bugs.
    `resample_filter` buffers keep their true value."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = {}
    for k in sorted(shapes):
        shp = tuple(shapes[k])
        if k.endswith("resample_filter"):
            out[k] = ops.setup_filter([1, 3, 3, 1])
            continue
        v = rng.standard_normal(shp).astype(np.float32)
        if k.endswith(".bias"):
            v = 0.1 * v + (1.0 if ".affine." in k else 0.0)
        elif k.endswith("noise_strength"):
            v = 0.1 + 0.05 * v
        elif k.endswith("w_avg"):
            v = 0.1 * v
        out[k] = torch.from_numpy(np.ascontiguousarray(v))
    return out


# ------------------------------------------------------------------------------------------------- loss phases (a23)
def run_G(g_sd: SD, cfg: StyleGANConfig, z: Tensor, c, h, style_mixing_prob: float, buffers_out: dict):
    """StyleGAN2Loss.run_G loss.py:57-76: training-mode mapping (w_avg tracked on the first call only), optional style
    mixing (the same three RNG calls in the same order), synthesis with its default noise_mode='random'."""
    kw = dict(z_dim=cfg.z_dim, num_layers=cfg.map_layers, num_ws=cfg.num_ws)
    ws = mapping(g_sd, "mapping", cfg, z, c, h, training=True, w_avg_beta=cfg.w_avg_beta, buffers_out=buffers_out, **kw)
    if style_mixing_prob > 0:
        cutoff = torch.empty([], dtype=torch.int64).random_(1, ws.shape[1])
        cutoff = torch.where(torch.rand([]) < style_mixing_prob, cutoff, torch.full_like(cutoff, ws.shape[1]))
        ws2 = mapping(g_sd, "mapping", cfg, torch.randn_like(z), c, h, **kw)  # skip_w_avg_update=True
        ws = torch.cat([ws[:, :int(cutoff)], ws2[:, int(cutoff):]], 1)
    img = synthesis(g_sd, "synthesis", cfg, ws, noise_mode="random")
    return img, ws


def accumulate_gradients(phase: str, g_sd: SD, d_sd: SD, cfg: StyleGANConfig, real_img, real_c, real_h, gen_z, gen_c,
                         gen_h, gain: float, *, pl_mean: Tensor, style_mixing_prob: float = 0.9, r1_gamma: float = 10.0,
                         pl_batch_shrink: int = 2, pl_decay: float = 0.01, pl_weight: float = 2.0) -> dict:
    """StyleGAN2Loss.accumulate_gradients loss.py:85-194 (no augment pipe).  Gradients accumulate into the `.grad` of
    whichever state-dict tensors have requires_grad (the caller toggles G / D as training_loop.py:395-425 does).
    Returns the scalar losses, the updated `pl_mean` and buffer updates (`mapping.w_avg`)."""
    assert phase in ["Gmain", "Greg", "Gboth", "Dmain", "Dreg", "Dboth"]
    do_Gmain, do_Dmain = phase in ["Gmain", "Gboth"], phase in ["Dmain", "Dboth"]
    do_Gpl = phase in ["Greg", "Gboth"] and pl_weight != 0
    do_Dr1 = phase in ["Dreg", "Dboth"] and r1_gamma != 0
    out = {"buffers": {}, "pl_mean": pl_mean}
    sp = torch.nn.functional.softplus
    if do_Gmain:  # :97-111
        img, _ = run_G(g_sd, cfg, gen_z, gen_c, gen_h, style_mixing_prob, out["buffers"])
        loss = sp(-discriminator(d_sd, cfg, img, gen_c, gen_h))
        out["loss_Gmain"] = loss.detach()
        loss.mean().mul(gain).backward()
    if do_Gpl:  # :113-143: path-length regulariser, a double backward through synthesis w.r.t. ws
        bs = gen_z.shape[0] // pl_batch_shrink
        img, ws = run_G(g_sd, cfg, gen_z[:bs], None if gen_c is None else gen_c[:bs], None if gen_h is None else gen_h[:bs],
                        style_mixing_prob, out["buffers"])
        pl_noise = torch.randn_like(img) / math.sqrt(img.shape[2] * img.shape[3])
        pl_grads = torch.autograd.grad([(img * pl_noise).sum()], [ws], create_graph=True, only_inputs=True)[0]
        pl_lengths = pl_grads.square().sum(2).mean(1).sqrt()
        new_mean = pl_mean.lerp(pl_lengths.mean(), pl_decay)
        out["pl_mean"] = new_mean.detach()
        pl_penalty = (pl_lengths - new_mean).square()
        loss_pl = pl_penalty * pl_weight
        out["loss_Gpl"] = loss_pl.detach()
        (img[:, 0, 0, 0] * 0 + loss_pl).mean().mul(gain).backward()
    loss_dgen = 0
    if do_Dmain:  # :146-160
        img, _ = run_G(g_sd, cfg, gen_z, gen_c, gen_h, style_mixing_prob, out["buffers"])
        loss_dgen = sp(discriminator(d_sd, cfg, img, gen_c, gen_h))
        out["loss_Dgen"] = loss_dgen.detach()
        loss_dgen.mean().mul(gain).backward()
    if do_Dmain or do_Dr1:  # :164-194
        real_tmp = real_img.detach().requires_grad_(do_Dr1)
        real_logits = discriminator(d_sd, cfg, real_tmp, real_c, real_h)
        loss_real = 0
        if do_Dmain:
            loss_real = sp(-real_logits)
            out["loss_Dreal"] = loss_real.detach()
        loss_r1 = 0
        if do_Dr1:
            r1_grads = torch.autograd.grad([real_logits.sum()], [real_tmp], create_graph=True, only_inputs=True)[0]
            r1_penalty = r1_grads.square().sum([1, 2, 3])
            loss_r1 = r1_penalty * (r1_gamma / 2)
            out["loss_Dr1"] = loss_r1.detach()
        (real_logits * 0 + loss_real + loss_r1).mean().mul(gain).backward()
    return out
