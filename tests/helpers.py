"""Shared test utilities: build the B200 modules with the oracle's synthetic weights, golden loading, smoke step."""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from oracle import biggan_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    with open(os.path.join(GOLD, f"biggan_{name}.json")) as f:
        meta = json.load(f)
    data = np.load(os.path.join(GOLD, f"biggan_{name}.npz"))
    cfg = O.BigGANConfig(**meta["config"])
    return cfg, meta, {k: torch.from_numpy(data[k]) for k in data.files}


def model_kwargs(cfg: O.BigGANConfig):
    return dict(G_ch=cfg.G_ch, D_ch=cfg.D_ch, dim_z=cfg.dim_z, resolution=cfg.resolution, G_attn=cfg.G_attn,
                D_attn=cfg.D_attn, n_classes=cfg.n_classes, G_shared=True, shared_dim=cfg.shared_dim, hier=cfg.hier,
                BN_eps=cfg.BN_eps, SN_eps=cfg.SN_eps, class_cond=cfg.class_cond, instance_cond=cfg.instance_cond,
                G_shared_feat=True, shared_dim_feat=cfg.shared_dim_feat, skip_init=True)


def make_models(cfg: O.BigGANConfig, device, compute_dtype, seed):
    """B200 Generator/Discriminator loaded (strict) with the same synthetic weights the golden vectors were made with."""
    from ic_gan_b200.biggan import Discriminator, Generator
    kw = model_kwargs(cfg)
    G = Generator(no_optim=True, compute_dtype=compute_dtype, **kw)
    D = Discriminator(embedded_optimizer=False, compute_dtype=compute_dtype, **kw)
    gs, ds = O.state_shapes(cfg)
    g_sd, d_sd = O.synth_state_dict(gs, seed), O.synth_state_dict(ds, seed + 1)
    G.load_state_dict(g_sd, strict=True)
    D.load_state_dict(d_sd, strict=True)
    return G.to(device), D.to(device), g_sd, d_sd


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def smoke_step():
    """One tiny G+D step (D phase + G phase, fp32 parity mode and bf16 tensor-core mode) on cuda:0 vs the CPU oracle."""
    from ic_gan_b200.biggan import G_D
    dev = torch.device("cuda", 0)
    cfg = O.BigGANConfig(resolution=32, G_ch=16, D_ch=16, G_attn="16", D_attn="16", n_classes=10, shared_dim=32,
                         shared_dim_feat=64, class_cond=True, instance_cond=True)
    B = 4
    g = torch.Generator().manual_seed(5)
    z = torch.randn(B, cfg.eff_dim_z, generator=g)
    feats = torch.nn.functional.normalize(torch.randn(B, 2048, generator=g), dim=1)
    feats_r = torch.nn.functional.normalize(torch.randn(B, 2048, generator=g), dim=1)
    x = torch.rand(B, 3, 32, 32, generator=g) * 2 - 1
    lab = torch.randint(0, 10, (B,), generator=g)
    lab_r = torch.randint(0, 10, (B,), generator=g)
    for cdt, tol in ((torch.float32, 2e-3), (torch.bfloat16, 0.25)):
        G, D, g_sd, d_sd = make_models(cfg, dev, cdt, 31)
        G.train(); D.train()
        GD = G_D(G, D)
        for p in G.parameters():
            p.requires_grad_(False)
        d_fake, d_real = GD(z.to(dev), lab.to(dev), feats.to(dev), x.to(dev), lab_r.to(dev), feats_r.to(dev))
        loss = torch.relu(1 - d_real).mean() + torch.relu(1 + d_fake).mean()
        loss.backward()
        for k, v in d_sd.items():
            if O.is_param(k, v):
                v.requires_grad_(True)
        o_fake, o_real = O.gd_forward(g_sd, d_sd, cfg, z, lab, feats, x, lab_r, feats_r, train_G=False)
        a, b = O.loss_hinge_dis(o_fake, o_real)
        (a + b).backward()
        err = (torch.cat([d_fake, d_real]).float().cpu() - torch.cat([o_fake, o_real])).abs().max().item()
        scale = torch.cat([o_fake, o_real]).abs().max().item()
        assert err <= tol * max(1.0, scale), f"smoke: D logits differ from the oracle by {err} ({cdt})"
        gw = D.blocks[1][0].conv1.weight.grad
        ge = rel_l2(gw, d_sd["blocks.1.0.conv1.weight"].grad)
        assert ge <= (2e-3 if cdt == torch.float32 else 0.2), f"smoke: conv weight grad rel-L2 {ge} ({cdt})"
        print(f"[smoke] {cdt}: D logits max-abs err {err:.3e}, conv1 wgrad rel-L2 {ge:.3e}")
    torch.cuda.synchronize()
