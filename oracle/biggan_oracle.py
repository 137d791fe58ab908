"""CPU oracle for IC-GAN's BigGAN hot path — TEST INFRASTRUCTURE ONLY.

A functional, float32, plain-PyTorch restatement of the reference algorithm (facebookresearch/ic_gan @ 8eff2f7),
operating directly on state_dicts that use the reference's key names and shapes.  Only tests/, bench.py's
cpu_baseline / --impl reference leg and __graft_entry__.smoke() may import this module; the product path
(ic_gan_b200/) never does.

Pinned against the live reference by oracle/make_golden.py (run in the build container, where /root/reference exists);
the resulting vectors live in tests/golden/ and tests/test_oracle_golden.py re-checks the oracle against them.

Reference lines restated here (paths relative to the reference repo):
  spectral norm ........ BigGAN_PyTorch/layers.py:39-61 (power_iteration), :98-112 (SN.W_)
  SNConv2d / SNLinear .. layers.py:144-153, :164-165       SNEmbedding .. layers.py:199-200
  Attention ............ layers.py:227-244                 ccbn / bn .... layers.py:398-437, :485-503
  GBlock / DBlock ...... layers.py:542-552, :587-613
  G_arch / D_arch ...... BigGAN_PyTorch/BigGAN.py:32-85, :390-432
  Generator.forward .... BigGAN.py:350-386                 Discriminator.forward .. BigGAN.py:617-642
  G_D.forward .......... BigGAN.py:655-711                 hinge losses .. BigGAN_PyTorch/losses.py:24-38
  training step ........ BigGAN_PyTorch/train_fns.py:40-191 ; EMA utils.py:1039-1067
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ----------------------------------------------------------------------------- configuration
@dataclass
class BigGANConfig:
    resolution: int = 64
    G_ch: int = 64
    D_ch: int = 64
    dim_z: int = 120
    bottom_width: int = 4
    G_attn: str = "32"
    D_attn: str = "32"
    n_classes: int = 1000
    shared_dim: int = 128
    shared_dim_feat: int = 512
    hier: bool = True
    class_cond: bool = False
    instance_cond: bool = True
    BN_eps: float = 1e-5
    SN_eps: float = 1e-6
    feat_dim: int = 2048

    # derived quantities (BigGAN.py:166-217)
    def g_arch(self):
        return g_arch(self.G_ch, self.G_attn, self.resolution)

    def d_arch(self):
        return d_arch(self.D_ch, self.D_attn, self.resolution)

    @property
    def num_slots(self):
        return len(self.g_arch()["in"]) + 1 if self.hier else 1

    @property
    def z_chunk(self):
        return self.dim_z // self.num_slots if self.hier else 0

    @property
    def eff_dim_z(self):  # dim_z is silently rounded down to a multiple of the chunk (BigGAN.py:172-177)
        return self.z_chunk * self.num_slots if self.hier else self.dim_z


_G_MULT = {
    512: ([16, 16, 8, 8, 4, 2, 1], [16, 8, 8, 4, 2, 1, 1]),
    256: ([16, 16, 8, 8, 4, 2], [16, 8, 8, 4, 2, 1]),
    128: ([16, 16, 8, 4, 2], [16, 8, 4, 2, 1]),
    64: ([16, 16, 8, 4], [16, 8, 4, 2]),
    32: ([4, 4, 4], [4, 4, 4]),
}
_D_MULT = {
    256: ([1, 2, 4, 8, 8, 16], [1, 2, 4, 8, 8, 16, 16], [128, 64, 32, 16, 8, 4, 4], 6),
    128: ([1, 2, 4, 8, 16], [1, 2, 4, 8, 16, 16], [64, 32, 16, 8, 4, 4], 5),
    64: ([1, 2, 4, 8], [1, 2, 4, 8, 16], [32, 16, 8, 4, 4], 4),
    32: ([4, 4, 4], [4, 4, 4, 4], [16, 16, 16, 16], 2),
}


def _attn_set(spec: str):
    return {int(s) for s in spec.split("_") if s}


def g_arch(ch: int, attention: str, resolution: int):
    mi, mo = _G_MULT[resolution]
    res = [8 * 2 ** i for i in range(len(mo))]
    att = _attn_set(attention)
    return {"in": [ch * m for m in mi], "out": [ch * m for m in mo], "res": res, "attn": [r in att for r in res]}


def d_arch(ch: int, attention: str, resolution: int):
    mi, mo, res, n_down = _D_MULT[resolution]
    att = _attn_set(attention)
    return {
        "in": [3] + [ch * m for m in mi],
        "out": [ch * m for m in mo],
        "down": [i < n_down for i in range(len(mo))],
        "res": res,
        "attn": [r in att for r in res],
    }


# ----------------------------------------------------------------------------- spectral norm
def _normalize(v: Tensor, eps: float) -> Tensor:
    return v / v.norm().clamp_min(eps)  # F.normalize on a [1, n] row


def sn_weight(sd: Dict[str, Tensor], prefix: str, training: bool, eps: float) -> Tensor:
    """W / sigma with ONE power iteration per call, also in eval (layers.py:98-112). u is updated in place only when
    training; sigma = v W^T u'^T is differentiated w.r.t. W with u', v held constant (layers.py:59)."""
    W = sd[prefix + "weight"]
    Wm = W.reshape(W.shape[0], -1)
    u = sd[prefix + "u0"]
    with torch.no_grad():
        v = _normalize(u @ Wm, eps)
        u_new = _normalize(v @ Wm.t(), eps)
        if training:
            u.copy_(u_new)
    sigma = (v @ Wm.t() @ u_new.t()).squeeze()
    if training:
        with torch.no_grad():
            sd[prefix + "sv0"].copy_(sigma.reshape(1))
    return W / sigma


def sn_conv(sd, prefix, x, training, eps, padding):
    return F.conv2d(x, sn_weight(sd, prefix, training, eps), sd.get(prefix + "bias"), 1, padding)


def sn_linear(sd, prefix, x, training, eps):
    return F.linear(x, sn_weight(sd, prefix, training, eps), sd.get(prefix + "bias"))


# ----------------------------------------------------------------------------- normalisation
class _BatchNormTrain(torch.autograd.Function):
    """Training-mode batch norm without affine, with the fused backward of ATen's native_batch_norm_backward
    (dx = invstd * (dy - mean(dy) - xhat * mean(dy * xhat))).  Differentiating mean/var through separate autograd ops is
    mathematically identical but loses ~2 digits in float32 to cancellation, which would un-pin the oracle."""

    @staticmethod
    def forward(ctx, x, eps):
        mean = x.mean(dim=(0, 2, 3))
        var = x.var(dim=(0, 2, 3), unbiased=False)
        invstd = torch.rsqrt(var + eps)
        xhat = (x - mean[None, :, None, None]) * invstd[None, :, None, None]
        ctx.save_for_backward(xhat, invstd)
        ctx.mark_non_differentiable(mean, var)
        return xhat, mean, var

    @staticmethod
    def backward(ctx, dy, _dm, _dv):
        xhat, invstd = ctx.saved_tensors
        m1 = dy.mean(dim=(0, 2, 3))[None, :, None, None]
        m2 = (dy * xhat).mean(dim=(0, 2, 3))[None, :, None, None]
        return (dy - m1 - xhat * m2) * invstd[None, :, None, None], None


def batch_norm(x: Tensor, sd, prefix: str, training: bool, eps: float, momentum: float = 0.1) -> Tensor:
    """F.batch_norm without affine: batch statistics (biased variance) when training, running buffers updated with the
    UNBIASED variance and momentum 0.1; stored statistics in eval (layers.py:412-421)."""
    mean_buf, var_buf = sd[prefix + "stored_mean"], sd[prefix + "stored_var"]
    if training:
        n = x.numel() // x.shape[1]
        xhat, mean, var = _BatchNormTrain.apply(x, eps)
        with torch.no_grad():
            mean_buf.mul_(1 - momentum).add_(momentum * mean)
            var_buf.mul_(1 - momentum).add_(momentum * var * (n / max(n - 1, 1)))
        return xhat
    mean, var = mean_buf, var_buf
    return (x - mean[None, :, None, None]) * torch.rsqrt(var[None, :, None, None] + eps)


def ccbn(sd, prefix, x, y, training, cfg: BigGANConfig):
    gain = 1 + sn_linear(sd, prefix + "gain.", y, training, cfg.SN_eps)
    bias = sn_linear(sd, prefix + "bias.", y, training, cfg.SN_eps)
    out = batch_norm(x, sd, prefix, training, cfg.BN_eps)
    return out * gain[:, :, None, None] + bias[:, :, None, None]


def plain_bn(sd, prefix, x, training, cfg: BigGANConfig):
    out = batch_norm(x, sd, prefix, training, cfg.BN_eps)
    return out * sd[prefix + "gain"][None, :, None, None] + sd[prefix + "bias"][None, :, None, None]


# ----------------------------------------------------------------------------- blocks
def attention(sd, prefix, x, training, eps):
    B, C, H, W = x.shape
    theta = sn_conv(sd, prefix + "theta.", x, training, eps, 0).reshape(B, C // 8, H * W)
    phi = F.max_pool2d(sn_conv(sd, prefix + "phi.", x, training, eps, 0), 2).reshape(B, C // 8, H * W // 4)
    g = F.max_pool2d(sn_conv(sd, prefix + "g.", x, training, eps, 0), 2).reshape(B, C // 2, H * W // 4)
    beta = torch.softmax(theta.transpose(1, 2) @ phi, dim=-1)
    o = (g @ beta.transpose(1, 2)).reshape(B, C // 2, H, W)
    return sd[prefix + "gamma"] * sn_conv(sd, prefix + "o.", o, training, eps, 0) + x


def g_block(sd, prefix, x, y, training, cfg):
    h = F.relu(ccbn(sd, prefix + "bn1.", x, y, training, cfg))
    h = F.interpolate(h, scale_factor=2)  # nearest (BigGAN.py:260)
    xs = F.interpolate(x, scale_factor=2)
    h = sn_conv(sd, prefix + "conv1.", h, training, cfg.SN_eps, 1)
    h = F.relu(ccbn(sd, prefix + "bn2.", h, y, training, cfg))
    h = sn_conv(sd, prefix + "conv2.", h, training, cfg.SN_eps, 1)
    return h + sn_conv(sd, prefix + "conv_sc.", xs, training, cfg.SN_eps, 0)


def d_block(sd, prefix, x, training, cfg, preact: bool, down: bool):
    eps = cfg.SN_eps
    h = F.relu(x) if preact else x
    h = sn_conv(sd, prefix + "conv1.", h, training, eps, 1)
    h = sn_conv(sd, prefix + "conv2.", F.relu(h), training, eps, 1)
    if down:
        h = F.avg_pool2d(h, 2)
    has_sc = (prefix + "conv_sc.weight") in sd
    s = x
    if preact:
        if has_sc:
            s = sn_conv(sd, prefix + "conv_sc.", s, training, eps, 0)
        if down:
            s = F.avg_pool2d(s, 2)
    else:
        if down:
            s = F.avg_pool2d(s, 2)
        if has_sc:
            s = sn_conv(sd, prefix + "conv_sc.", s, training, eps, 0)
    return h + s


def g_block_deep(sd, prefix, x, y, training, cfg, out_channels: int, upsample: bool):
    """BigGANdeep.GBlock.forward (BigGANdeep.py:67-85): 1x1 down, BN-ReLU, (x2 nearest), 3x3, 3x3, 1x1 up; the shortcut
    drops channels instead of projecting them."""
    h = sn_conv(sd, prefix + "conv1.", F.relu(ccbn(sd, prefix + "bn1.", x, y, training, cfg)), training, cfg.SN_eps, 0)
    h = F.relu(ccbn(sd, prefix + "bn2.", h, y, training, cfg))
    if x.shape[1] != out_channels:
        x = x[:, :out_channels]
    if upsample:
        h, x = F.interpolate(h, scale_factor=2), F.interpolate(x, scale_factor=2)
    h = sn_conv(sd, prefix + "conv2.", h, training, cfg.SN_eps, 1)
    h = sn_conv(sd, prefix + "conv3.", F.relu(ccbn(sd, prefix + "bn3.", h, y, training, cfg)), training, cfg.SN_eps, 1)
    h = sn_conv(sd, prefix + "conv4.", F.relu(ccbn(sd, prefix + "bn4.", h, y, training, cfg)), training, cfg.SN_eps, 0)
    return h + x


def d_block_deep(sd, prefix, x, training, cfg, down: bool):
    """BigGANdeep.DBlock.forward (BigGANdeep.py:431-451): relu-1x1, relu-3x3, relu-3x3, relu, (avg-pool), 1x1; shortcut =
    (avg-pool) then concat with conv_sc when the channel count grows."""
    eps = cfg.SN_eps
    h = sn_conv(sd, prefix + "conv1.", F.relu(x), training, eps, 0)
    h = sn_conv(sd, prefix + "conv2.", F.relu(h), training, eps, 1)
    h = F.relu(sn_conv(sd, prefix + "conv3.", F.relu(h), training, eps, 1))
    if down:
        h = F.avg_pool2d(h, 2)
    h = sn_conv(sd, prefix + "conv4.", h, training, eps, 0)
    s = F.avg_pool2d(x, 2) if down else x
    if (prefix + "conv_sc.weight") in sd:
        s = torch.cat([s, sn_conv(sd, prefix + "conv_sc.", s, training, eps, 0)], 1)
    return h + s


# ----------------------------------------------------------------------------- networks
def generator_forward(sd, cfg: BigGANConfig, z: Tensor, label: Optional[Tensor], feats: Optional[Tensor],
                      training: bool) -> Tensor:
    arch = cfg.g_arch()
    emb = []
    if label is not None:
        emb.append(sd["shared.weight"][label])  # plain nn.Embedding, never SN (BigGAN.py:202-204)
    if feats is not None:
        emb.append(sn_linear(sd, "shared_feat.", feats, training, cfg.SN_eps))
    y = torch.cat(emb, dim=-1)
    if cfg.hier:
        zs = torch.split(z, cfg.z_chunk, 1)
        z0 = zs[0]
        ys = [torch.cat([y, zc], 1) for zc in zs[1:]]
    else:
        z0, ys = z, [y] * len(arch["out"])
    h = sn_linear(sd, "linear.", z0, training, cfg.SN_eps)
    h = h.reshape(h.shape[0], -1, cfg.bottom_width, cfg.bottom_width)
    for i in range(len(arch["out"])):
        h = g_block(sd, f"blocks.{i}.0.", h, ys[i], training, cfg)
        if arch["attn"][i]:
            h = attention(sd, f"blocks.{i}.1.", h, training, cfg.SN_eps)
    h = F.relu(plain_bn(sd, "output_layer.0.", h, training, cfg))
    return torch.tanh(sn_conv(sd, "output_layer.2.", h, training, cfg.SN_eps, 1))


def discriminator_forward(sd, cfg: BigGANConfig, x: Tensor, y: Optional[Tensor], feat: Optional[Tensor],
                          training: bool) -> Tensor:
    arch = cfg.d_arch()
    h = x
    for i in range(len(arch["out"])):
        h = d_block(sd, f"blocks.{i}.0.", h, training, cfg, preact=i > 0, down=arch["down"][i])
        if arch["attn"][i]:
            h = attention(sd, f"blocks.{i}.1.", h, training, cfg.SN_eps)
    h = F.relu(h).sum(dim=(2, 3))
    out = sn_linear(sd, "linear.", h, training, cfg.SN_eps)
    proj = []
    if y is not None:
        proj.append(F.embedding(y, sn_weight(sd, "embed.", training, cfg.SN_eps)))
    if feat is not None:
        proj.append(sn_linear(sd, "linear_feat.", feat, training, cfg.SN_eps))
    if proj:
        out = out + (torch.cat(proj, dim=-1) * h).sum(1, keepdim=True)
    return out


def gd_forward(g_sd, d_sd, cfg, z, gy, feats_g, x=None, dy=None, feats=None, train_G=False, training=True):
    """G_D.forward with split_D=False: one D pass over cat(fake, real) (BigGAN.py:693-706)."""
    with torch.set_grad_enabled(train_G):
        g_z = generator_forward(g_sd, cfg, z, gy, feats_g, training)
    d_in = torch.cat([g_z, x], 0) if x is not None else g_z
    d_cls = None if gy is None else (torch.cat([gy, dy], 0) if dy is not None else gy)
    d_feat = None if feats_g is None else (torch.cat([feats_g, feats], 0) if feats is not None else feats_g)
    out = discriminator_forward(d_sd, cfg, d_in, d_cls, d_feat, training)
    if x is not None:
        return torch.split(out, [g_z.shape[0], x.shape[0]])
    return out


def loss_hinge_dis(d_fake, d_real):
    return F.relu(1.0 - d_real).mean(), F.relu(1.0 + d_fake).mean()


def loss_hinge_gen(d_fake):
    return -d_fake.mean()


# ----------------------------------------------------------------------------- training step
PARAM_SUFFIXES = ("weight", "bias", "gamma", "gain")


def is_param(key: str, tensor: Tensor) -> bool:
    """Parameters vs buffers in a reference-layout state_dict (buffers: u0, sv0, stored_mean, stored_var)."""
    leaf = key.rsplit(".", 1)[-1]
    return leaf in PARAM_SUFFIXES and tensor.dtype.is_floating_point


@dataclass
class StepState:
    g_sd: Dict[str, Tensor]
    d_sd: Dict[str, Tensor]
    ema_sd: Optional[Dict[str, Tensor]] = None
    opt_g: Optional[torch.optim.Optimizer] = None
    opt_d: Optional[torch.optim.Optimizer] = None
    itr: int = 0
    extra: dict = field(default_factory=dict)


def make_step_state(g_sd, d_sd, G_lr=5e-5, D_lr=2e-4, B1=0.0, B2=0.999, adam_eps=1e-6, ema=True) -> StepState:
    for sd in (g_sd, d_sd):
        for k, v in sd.items():
            if is_param(k, v):
                v.requires_grad_(True)
    gp = [v for k, v in g_sd.items() if is_param(k, v)]
    dp = [v for k, v in d_sd.items() if is_param(k, v)]
    st = StepState(g_sd, d_sd)
    st.opt_g = torch.optim.Adam(gp, lr=G_lr, betas=(B1, B2), weight_decay=0, eps=adam_eps)
    st.opt_d = torch.optim.Adam(dp, lr=D_lr, betas=(B1, B2), weight_decay=0, eps=adam_eps)
    if ema:
        st.ema_sd = {k: v.detach().clone() for k, v in g_sd.items()}
    return st


def train_step(st: StepState, cfg: BigGANConfig, x: Tensor, y: Optional[Tensor], feats: Optional[Tensor],
               sample_cond, batch_size: int, num_D_steps=1, num_D_acc=1, num_G_acc=1, ema_decay=0.9999,
               ema_start=0) -> dict:
    """One GAN_training_function.train call (train_fns.py:40-191), toggle_grads=True, hinge loss, no ortho reg.
    sample_cond() -> (z, labels_or_None, feats_or_None) of at least batch_size rows."""
    st.opt_g.zero_grad(set_to_none=False)
    st.opt_d.zero_grad(set_to_none=False)
    xs = torch.split(x, batch_size)
    ys = torch.split(y, batch_size) if y is not None else None
    fs = torch.split(feats, batch_size) if feats is not None else None
    g_params = [v for k, v in st.g_sd.items() if is_param(k, v)]
    d_params = [v for k, v in st.d_sd.items() if is_param(k, v)]
    for p in d_params:
        p.requires_grad_(True)
    for p in g_params:
        p.requires_grad_(False)
    counter = 0
    for _ in range(num_D_steps):
        st.opt_d.zero_grad(set_to_none=False)
        for _ in range(num_D_acc):
            z_, lab_g, f_g = sample_cond()
            z_ = z_[:batch_size]
            lab_g = None if lab_g is None else lab_g[:batch_size].long()
            f_g = None if f_g is None else f_g[:batch_size]
            d_fake, d_real = gd_forward(st.g_sd, st.d_sd, cfg, z_, lab_g, f_g, xs[counter],
                                        None if ys is None else ys[counter], None if fs is None else fs[counter],
                                        train_G=False)
            l_real, l_fake = loss_hinge_dis(d_fake, d_real)
            ((l_real + l_fake) / float(num_D_acc)).backward()
            counter += 1
        st.opt_d.step()
    for p in d_params:
        p.requires_grad_(False)
    for p in g_params:
        p.requires_grad_(True)
    st.opt_g.zero_grad(set_to_none=False)
    for _ in range(num_G_acc):
        z_, lab_g, f_g = sample_cond()
        lab_g = None if lab_g is None else lab_g.long()
        d_fake = gd_forward(st.g_sd, st.d_sd, cfg, z_, lab_g, f_g, train_G=True)
        g_loss = loss_hinge_gen(d_fake) / float(num_G_acc)
        g_loss.backward()
    st.opt_g.step()
    if st.ema_sd is not None:  # utils.py:1055-1067: every state_dict entry, decay 0 before ema_start
        decay = 0.0 if (st.itr and st.itr < ema_start) else ema_decay  # `if itr and itr < start_itr` (utils.py:1058)
        with torch.no_grad():
            for k, v in st.g_sd.items():
                st.ema_sd[k].copy_(st.ema_sd[k] * decay + v.detach() * (1 - decay))
    st.itr += 1
    return {"G_loss": g_loss.item(), "D_loss_real": l_real.item(), "D_loss_fake": l_fake.item()}


# ----------------------------------------------------------------------------- synthetic weights
def synth_state_dict(shapes: Dict[str, List[int]], seed: int) -> Dict[str, Tensor]:
    """Deterministic, platform-independent weights for a reference-layout state_dict (numpy PCG64 streams keyed by the
    sorted key order). Used instead of storing multi-megabyte checkpoints in tests/golden/."""
    import numpy as np

    out = {}
    for i, key in enumerate(sorted(shapes)):
        shape = tuple(shapes[key])
        rng = np.random.default_rng([seed, i])
        leaf = key.rsplit(".", 1)[-1]
        parent = key.rsplit(".", 2)[-2] if key.count(".") >= 1 else ""
        if leaf == "weight":
            if parent in ("shared", "embed"):
                a = rng.standard_normal(shape) * 0.5
            else:
                fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else shape[0]
                a = rng.standard_normal(shape) / math.sqrt(fan_in)
        elif leaf == "bias":
            a = rng.standard_normal(shape) * 0.1
        elif leaf == "gain":  # plain bn gain (output_layer.0.gain)
            a = 1.0 + 0.1 * rng.standard_normal(shape)
        elif leaf == "gamma":
            a = np.asarray(0.7)
        elif leaf == "stored_mean":
            a = rng.standard_normal(shape) * 0.1
        elif leaf == "stored_var":
            a = rng.uniform(0.5, 1.5, shape)
        elif leaf.startswith("u"):
            a = rng.standard_normal(shape)
        elif leaf.startswith("sv"):
            a = np.ones(shape)
        else:
            raise KeyError(f"unexpected state_dict key {key}")
        out[key] = torch.from_numpy(np.asarray(a, dtype=np.float32).reshape(shape)).clone()
    return out


def state_shapes(cfg: BigGANConfig):
    """Key -> shape tables of the reference G and D state_dicts for cfg (verified against the live reference in
    oracle/make_golden.py). Lets tests build weights without the reference present."""
    g, d = {}, {}

    def sn(tab, p, w_shape, bias=True, n_out=None):
        tab[p + "weight"] = list(w_shape)
        if bias:
            tab[p + "bias"] = [w_shape[0]]
        tab[p + "u0"] = [1, n_out if n_out is not None else w_shape[0]]
        tab[p + "sv0"] = [1]

    def attn(tab, p, C):
        sn(tab, p + "theta.", (C // 8, C, 1, 1), bias=False)
        sn(tab, p + "phi.", (C // 8, C, 1, 1), bias=False)
        sn(tab, p + "g.", (C // 2, C, 1, 1), bias=False)
        sn(tab, p + "o.", (C, C // 2, 1, 1), bias=False)
        tab[p + "gamma"] = []

    ga = cfg.g_arch()
    cond = cfg.z_chunk + (cfg.shared_dim if cfg.class_cond else 0) + (cfg.shared_dim_feat if cfg.instance_cond else 0)
    g["shared.weight"] = [cfg.n_classes, cfg.shared_dim]
    sn(g, "shared_feat.", (cfg.shared_dim_feat, cfg.feat_dim))
    sn(g, "linear.", (ga["in"][0] * cfg.bottom_width ** 2, cfg.eff_dim_z // cfg.num_slots))
    for i, (ci, co) in enumerate(zip(ga["in"], ga["out"])):
        p = f"blocks.{i}.0."
        sn(g, p + "conv1.", (co, ci, 3, 3))
        sn(g, p + "conv2.", (co, co, 3, 3))
        sn(g, p + "conv_sc.", (co, ci, 1, 1))
        for name, c in (("bn1.", ci), ("bn2.", co)):
            sn(g, p + name + "gain.", (c, cond), bias=False)
            sn(g, p + name + "bias.", (c, cond), bias=False)
            g[p + name + "stored_mean"] = [c]
            g[p + name + "stored_var"] = [c]
        if ga["attn"][i]:
            attn(g, f"blocks.{i}.1.", co)
    c_last = ga["out"][-1]
    for k in ("gain", "bias", "stored_mean", "stored_var"):
        g["output_layer.0." + k] = [c_last]
    sn(g, "output_layer.2.", (3, c_last, 3, 3))

    da = cfg.d_arch()
    for i, (ci, co) in enumerate(zip(da["in"], da["out"])):
        p = f"blocks.{i}.0."
        sn(d, p + "conv1.", (co, ci, 3, 3))
        sn(d, p + "conv2.", (co, co, 3, 3))
        if ci != co or da["down"][i]:
            sn(d, p + "conv_sc.", (co, ci, 1, 1))
        if da["attn"][i]:
            attn(d, f"blocks.{i}.1.", co)
    c_last = da["out"][-1]
    sn(d, "linear.", (1, c_last))
    if cfg.class_cond and cfg.instance_cond:
        sn(d, "linear_feat.", (c_last // 2, cfg.feat_dim))
        sn(d, "embed.", (cfg.n_classes, c_last // 2), bias=False, n_out=cfg.n_classes)
    elif cfg.class_cond:
        sn(d, "embed.", (cfg.n_classes, c_last), bias=False, n_out=cfg.n_classes)
    elif cfg.instance_cond:
        sn(d, "linear_feat.", (c_last, cfg.feat_dim))
    return g, d
