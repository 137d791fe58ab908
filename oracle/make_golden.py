"""Pin the oracle against the LIVE reference and freeze golden vectors into tests/golden/.

Run in the build container only (needs /root/reference, read-only):   python oracle/make_golden.py
The reference ships no golden vectors or tests for this path, so every fixture is produced here by
importing the reference's own modules on CPU/fp32, loading deterministic synthetic weights (oracle.synth_state_dict),
and recording inputs + outputs.  The script asserts oracle == reference before writing anything.
"""
from __future__ import annotations

import io
import contextlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("ICGAN_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")

from oracle import biggan_oracle as O  # noqa: E402


def _import_reference():
    sys.path.insert(0, REF)
    with contextlib.redirect_stdout(io.StringIO()):
        import BigGAN_PyTorch.BigGAN as RB  # noqa
        import BigGAN_PyTorch.losses as RL  # noqa
    return RB, RL


def _ref_models(RB, cfg: O.BigGANConfig):
    common = dict(resolution=cfg.resolution, n_classes=cfg.n_classes, SN_eps=cfg.SN_eps, class_cond=cfg.class_cond,
                  instance_cond=cfg.instance_cond, skip_init=True)
    with contextlib.redirect_stdout(io.StringIO()):
        G = RB.Generator(G_ch=cfg.G_ch, dim_z=cfg.dim_z, G_attn=cfg.G_attn, G_shared=True, shared_dim=cfg.shared_dim,
                         hier=cfg.hier, BN_eps=cfg.BN_eps, G_shared_feat=True, shared_dim_feat=cfg.shared_dim_feat,
                         no_optim=True, **common)
        D = RB.Discriminator(D_ch=cfg.D_ch, D_attn=cfg.D_attn, embedded_optimizer=False, **common)
    return G, D


def _check_shapes(G, D, cfg):
    gs, ds = O.state_shapes(cfg)
    for name, mod, tab in (("G", G, gs), ("D", D, ds)):
        ref = {k: list(v.shape) for k, v in mod.state_dict().items()}
        assert ref == tab, f"{name} state_dict layout differs: " + str(
            {k: (ref.get(k), tab.get(k)) for k in set(ref) ^ set(tab) | {k for k in ref if ref.get(k) != tab.get(k)}})
    return gs, ds


def _grad_digest(named_grads):
    out = {}
    for k, g in named_grads.items():
        g = g.detach().double().flatten()
        out[k] = [float(g.sum()), float(g.abs().sum()), float((g * torch.arange(1, g.numel() + 1).double().fmod(7.0)).sum())]
    return out


def _inputs(cfg, B, seed):
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(B, cfg.eff_dim_z, generator=g)
    feats = torch.nn.functional.normalize(torch.randn(B, cfg.feat_dim, generator=g), dim=1)
    feats_r = torch.nn.functional.normalize(torch.randn(B, cfg.feat_dim, generator=g), dim=1)
    x = torch.rand(B, 3, cfg.resolution, cfg.resolution, generator=g) * 2 - 1
    lab = torch.randint(0, cfg.n_classes, (B,), generator=g) if cfg.class_cond else None
    lab_r = torch.randint(0, cfg.n_classes, (B,), generator=g) if cfg.class_cond else None
    return z, feats, feats_r, x, lab, lab_r


def close(a, b, tol, what):
    err = (a.detach() - b.detach()).abs().max().item()
    assert err <= tol, f"{what}: oracle vs reference max-abs {err:.3e} > {tol}"
    return err


def biggan_case(RB, RL, name, cfg: O.BigGANConfig, B, seed):
    G, D = _ref_models(RB, cfg)
    gs, ds = _check_shapes(G, D, cfg)
    g_sd0, d_sd0 = O.synth_state_dict(gs, seed), O.synth_state_dict(ds, seed + 1)
    z, feats, feats_r, x, lab, lab_r = _inputs(cfg, B, seed + 2)
    fx = {"z": z, "feats_g": feats, "feats_r": feats_r, "x": x}
    if lab is not None:
        fx["label_g"], fx["label_r"] = lab, lab_r
    report = {}

    def fresh():
        G.load_state_dict({k: v.clone() for k, v in g_sd0.items()})
        D.load_state_dict({k: v.clone() for k, v in d_sd0.items()})
        return {k: v.clone() for k, v in g_sd0.items()}, {k: v.clone() for k, v in d_sd0.items()}

    # 1. G eval forward
    g_sd, d_sd = fresh()
    G.eval()
    with torch.no_grad():
        ref = G(z, lab, feats)
        mine = O.generator_forward(g_sd, cfg, z, lab, feats, training=False)
    report["G_eval"] = close(mine, ref, 2e-5, "G eval")
    fx["G_eval_out"] = ref

    # 2. D phase: G no-grad train-mode forward, D on cat(fake, real), hinge loss, backward
    g_sd, d_sd = fresh()
    G.train(); D.train()
    GD = RB.G_D(G, D)
    for p in G.parameters():
        p.requires_grad_(False)
    d_fake, d_real = GD(z, lab, feats, x, lab_r, feats_r, train_G=False)
    lr_, lf_ = RL.loss_hinge_dis(d_fake, d_real)
    (lr_ + lf_).backward()
    for k, v in d_sd.items():
        if O.is_param(k, v):
            v.requires_grad_(True)
    o_fake, o_real = O.gd_forward(g_sd, d_sd, cfg, z, lab, feats, x, lab_r, feats_r, train_G=False)
    a, b = O.loss_hinge_dis(o_fake, o_real)
    (a + b).backward()
    report["D_out"] = close(torch.cat([o_fake, o_real]), torch.cat([d_fake, d_real]), 5e-4, "D out")
    ref_dg = {k: p.grad for k, p in D.named_parameters()}
    for k, p in ref_dg.items():
        scale = max(1.0, p.abs().max().item())
        close(d_sd[k].grad / scale, p / scale, 2e-4, f"D grad {k}")
    for k, v in G.state_dict().items():  # buffers after a training forward (u0, sv0, BN running stats)
        if not O.is_param(k, v):
            close(g_sd[k], v, 1e-4, f"G buffer {k}")
    for k, v in D.state_dict().items():
        if not O.is_param(k, v):
            close(d_sd[k], v, 1e-4, f"D buffer {k}")
    fx["D_fake"], fx["D_real"] = d_fake.detach(), d_real.detach()
    fx["D_loss"] = torch.stack([lr_.detach(), lf_.detach()])
    digest = {"D_phase": _grad_digest(ref_dg)}
    bufs = {"G." + k: v.clone() for k, v in G.state_dict().items() if not O.is_param(k, v)}
    bufs.update({"D." + k: v.clone() for k, v in D.state_dict().items() if not O.is_param(k, v)})
    for k, v in bufs.items():
        fx["buf_after_Dphase/" + k] = v
    # a few complete gradients (small tensors) for elementwise comparison
    for k in list(ref_dg)[:3] + list(ref_dg)[-3:]:
        if ref_dg[k].numel() <= 70000:
            fx["D_grad/" + k] = ref_dg[k].clone()

    # 3. G phase: grads w.r.t. G through a frozen D
    g_sd, d_sd = fresh()
    G.train(); D.train()
    for p in G.parameters():
        p.requires_grad_(True)
        p.grad = None
    for p in D.parameters():
        p.requires_grad_(False)
    d_fake = GD(z, lab, feats, train_G=True)
    gl = RL.loss_hinge_gen(d_fake)
    gl.backward()
    for k, v in g_sd.items():
        if O.is_param(k, v):
            v.requires_grad_(True)
    o = O.gd_forward(g_sd, d_sd, cfg, z, lab, feats, train_G=True)
    O.loss_hinge_gen(o).backward()
    report["G_phase_out"] = close(o, d_fake, 5e-4, "G-phase D out")
    ref_gg = {k: p.grad for k, p in G.named_parameters() if p.grad is not None}
    for k, p in ref_gg.items():
        scale = max(1.0, p.abs().max().item())
        close(g_sd[k].grad / scale, p / scale, 2e-4, f"G grad {k}")
    digest["G_phase"] = _grad_digest(ref_gg)
    fx["G_phase_D_fake"] = d_fake.detach()
    for k in list(ref_gg)[:3] + list(ref_gg)[-4:]:
        if ref_gg[k].numel() <= 70000:
            fx["G_grad/" + k] = ref_gg[k].clone()

    np.savez_compressed(os.path.join(GOLD, f"biggan_{name}.npz"), **{k: v.detach().numpy() for k, v in fx.items()})
    meta = {"config": cfg.__dict__, "batch": B, "seed": seed, "grad_digest": digest, "oracle_vs_reference": report,
            "g_shapes": gs, "d_shapes": ds}
    with open(os.path.join(GOLD, f"biggan_{name}.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print(f"[golden] biggan_{name}: oracle==reference {report}")


def biggan_config1(RB):
    """BASELINE.json configs[0]: ic-64 generator forward, batch 8, 1000 stored instance features, CPU only."""
    cfg = O.BigGANConfig(resolution=64, G_ch=64, D_ch=64, G_attn="32", D_attn="32", class_cond=False,
                         instance_cond=True)
    G, D = _ref_models(RB, cfg)
    gs, _ = _check_shapes(G, D, cfg)
    sd = O.synth_state_dict(gs, 100)
    G.load_state_dict({k: v.clone() for k, v in sd.items()})
    G.eval()
    table = torch.nn.functional.normalize(torch.randn(1000, 2048, generator=torch.Generator().manual_seed(1)), dim=1)
    idx = torch.arange(0, 1000, 125)
    z = torch.randn(8, cfg.eff_dim_z, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        ref = G(z, None, table[idx])
        mine = O.generator_forward(sd, cfg, z, None, table[idx], training=False)
    err = close(mine, ref, 2e-5, "config1 G eval")
    np.savez_compressed(os.path.join(GOLD, "biggan_config1_ic64.npz"), z=z.numpy(), idx=idx.numpy(),
                        out=ref.numpy().astype(np.float16), out_sample=ref[:, :, ::8, ::8].numpy())
    with open(os.path.join(GOLD, "biggan_config1_ic64.json"), "w") as f:
        json.dump({"config": cfg.__dict__, "seed_weights": 100, "seed_table": 1, "seed_z": 0,
                   "oracle_vs_reference": err}, f, indent=1, sort_keys=True)
    print(f"[golden] biggan_config1_ic64: oracle==reference {err:.2e}")


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(8)
    RB, RL = _import_reference()
    biggan_case(RB, RL, "ic64_tiny", O.BigGANConfig(resolution=64, G_ch=16, D_ch=16, G_attn="32", D_attn="32",
                                                   class_cond=False, instance_cond=True), B=4, seed=12)
    biggan_case(RB, RL, "cc32_tiny", O.BigGANConfig(resolution=32, G_ch=16, D_ch=16, G_attn="16", D_attn="16",
                                                   n_classes=10, shared_dim=32, shared_dim_feat=64,
                                                   class_cond=True, instance_cond=True), B=4, seed=21)
    biggan_config1(RB)
    if "--all" in sys.argv or True:
        try:
            from oracle import make_golden_extra
            make_golden_extra.main(REF, GOLD)
        except ImportError:
            pass


if __name__ == "__main__":
    main()
