"""The StyleGAN2-ADA network oracle (oracle/stylegan_nets_oracle.py: mapping / synthesis / generator / discriminator,
SURVEY.md section 8 rows a21-a22) against golden vectors recorded from the live reference on the CPU
(oracle/make_golden_stylegan_nets.py).  CPU-only: this pins the oracle the B200 modules will be built against."""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import stylegan_nets_oracle as O
from oracle.make_golden_stylegan_nets import inputs

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 2e-5


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(GOLD, "stylegan_nets.json")) as f:
        meta = json.load(f)
    fx = dict(np.load(os.path.join(GOLD, "stylegan_nets.npz")))
    cfg = O.StyleGANConfig(**meta["cfg"])
    g_sd = O.synth_state_dict(meta["g_shapes"], meta["g_seed"])
    d_sd = O.synth_state_dict(meta["d_shapes"], meta["d_seed"])
    return cfg, meta, fx, g_sd, d_sd


def _close(got, ref, what, tol=TOL):
    got = got.detach().numpy() if isinstance(got, torch.Tensor) else got
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    err = np.abs(got - ref).max()
    scale = max(1.0, np.abs(ref).max())
    assert err <= tol * scale, f"{what}: max-abs err {err:.3e} (scale {scale:.3g})"


def test_mapping_and_truncation(gold):
    cfg, meta, fx, g_sd, _ = gold
    z, h, _ = inputs()
    kw = dict(z_dim=cfg.z_dim, num_layers=cfg.map_layers, num_ws=cfg.num_ws)
    _close(O.mapping(g_sd, "mapping", cfg, z, None, h, **kw), fx["ws"], "ws")
    _close(O.mapping(g_sd, "mapping", cfg, z, None, h, truncation_psi=0.7, truncation_cutoff=3, **kw), fx["ws_trunc"],
           "ws truncated")
    assert fx["ws"].shape == (4, cfg.num_ws, cfg.w_dim)


@pytest.mark.parametrize("mode,key", [("const", "img_const"), ("none", "img_none")])
def test_generator_images(gold, mode, key):
    cfg, meta, fx, g_sd, _ = gold
    z, h, _ = inputs()
    _close(O.generator(g_sd, cfg, z, None, h, noise_mode=mode), fx[key], key)


def test_generator_truncated_random_noise_and_training_mode(gold):
    cfg, meta, fx, g_sd, _ = gold
    z, h, _ = inputs()
    _close(O.generator(g_sd, cfg, z, None, h, truncation_psi=0.5), fx["img_trunc"], "img_trunc")
    torch.manual_seed(5)  # same global-stream randn calls, layer by layer, as the reference
    _close(O.generator(g_sd, cfg, z, None, h, noise_mode="random"), fx["img_random"], "img_random")
    bufs = {}
    img = O.generator(g_sd, cfg, z, None, h, training=True, buffers_out=bufs)
    _close(img, fx["img_train"], "img (training mode, non-fused modconv in the reference)")
    _close(bufs["mapping.w_avg"], fx["w_avg_after"], "w_avg after one training forward")


def test_discriminator_logits(gold):
    cfg, meta, fx, _, d_sd = gold
    _, h, x = inputs()
    _close(O.discriminator(d_sd, cfg, x, None, h), fx["d_real"], "D(real)", tol=2e-5)
    _close(O.discriminator(d_sd, cfg, torch.from_numpy(fx["img_const"]), None, h), fx["d_fake"], "D(fake)", tol=2e-5)


def test_loss_gradients(gold):
    """Non-saturating logistic G and D losses (training/loss.py:96-100, :126-150): parameter gradients through the
    whole oracle equal the reference's."""
    cfg, meta, fx, g_sd, d_sd = gold
    z, h, x = inputs()
    for k in meta["grad_keys_g"]:
        g_sd[k].requires_grad_(True)
    for k in meta["grad_keys_d"]:
        d_sd[k].requires_grad_(True)
    img = O.generator(g_sd, cfg, z, None, h, training=True, buffers_out={})
    loss_g = F.softplus(-O.discriminator(d_sd, cfg, img, None, h)).mean()
    assert abs(loss_g.item() - float(fx["loss_g"][0])) <= 1e-5
    grads = torch.autograd.grad(loss_g, [g_sd[k] for k in meta["grad_keys_g"]])
    for k, g in zip(meta["grad_keys_g"], grads):
        ref = fx["G_grad/" + k]
        rel = np.linalg.norm(g.numpy() - ref) / max(np.linalg.norm(ref), 1e-12)
        assert rel <= 2e-4, f"G grad {k}: rel-L2 {rel:.3e}"
    loss_d = F.softplus(O.discriminator(d_sd, cfg, img.detach(), None, h)).mean() + \
        F.softplus(-O.discriminator(d_sd, cfg, x, None, h)).mean()
    assert abs(loss_d.item() - float(fx["loss_d"][0])) <= 1e-5
    grads = torch.autograd.grad(loss_d, [d_sd[k] for k in meta["grad_keys_d"]])
    for k, g in zip(meta["grad_keys_d"], grads):
        ref = fx["D_grad/" + k]
        rel = np.linalg.norm(g.numpy() - ref) / max(np.linalg.norm(ref), 1e-12)
        assert rel <= 2e-4, f"D grad {k}: rel-L2 {rel:.3e}"


def test_minibatch_std_groups():
    """MinibatchStdLayer networks.py:906-927: batch 8, group 4 -> two groups, one extra channel constant per group."""
    g = torch.Generator().manual_seed(0)
    x = torch.randn(8, 6, 4, 4, generator=g)
    y = O.minibatch_std(x, 4, 1)
    assert y.shape == (8, 7, 4, 4) and torch.equal(y[:, :6], x)
    extra = y[:, 6]
    for n in range(8):
        assert torch.all(extra[n] == extra[n, 0, 0])
    assert torch.equal(extra[0], extra[2]) and torch.equal(extra[1], extra[3]) and not torch.equal(extra[0], extra[1])


@pytest.mark.parametrize("phase,seed,gain", [("Gmain", 31, 1.0), ("Greg", 32, 4.0), ("Dmain", 33, 1.0), ("Dreg", 34, 16.0)])
def test_loss_phases_match_reference(gold, phase, seed, gain):
    """StyleGAN2Loss.accumulate_gradients (training/loss.py:85-194) phase by phase: style mixing, random noise, the
    path-length regulariser (double backward through synthesis) and R1 (double backward through D), against golden
    gradients from the live reference with the same torch seed."""
    cfg, meta, _, _, _ = gold
    fx = dict(np.load(os.path.join(GOLD, "stylegan_loss.npz")))
    g_sd = O.synth_state_dict(meta["g_shapes"], meta["g_seed"])
    d_sd = O.synth_state_dict(meta["d_shapes"], meta["d_seed"])
    g_sd["mapping.w_avg"] = torch.from_numpy(fx[f"{phase}/w_avg_before"])
    pl_before = {"Gmain": "pl_mean_before", "Greg": "pl_mean_before"}.get(phase)
    pl_mean = torch.tensor(float(fx[pl_before][0] if pl_before else fx["Greg/pl_mean_after"][0]))
    is_g = phase.startswith("G")
    for k, v in (g_sd if is_g else d_sd).items():
        if v.dtype.is_floating_point and not k.endswith(("resample_filter", "noise_const", "w_avg")):
            v.requires_grad_(True)
    z, h, x = inputs()
    torch.manual_seed(seed)
    out = O.accumulate_gradients(phase, g_sd, d_sd, cfg, x, None, h, z, None, h, gain, pl_mean=pl_mean)
    keys = meta["grad_keys_g"] if is_g else meta["grad_keys_d"]
    sd = g_sd if is_g else d_sd
    for k in keys:
        ref = fx[f"{phase}/grad/{k}"]
        got = sd[k].grad.numpy() if sd[k].grad is not None else np.zeros_like(ref)
        denom = max(np.linalg.norm(ref), 1e-30)
        rel = np.linalg.norm(got - ref) / denom
        assert rel <= 2e-3, f"{phase} grad {k}: rel-L2 {rel:.3e} (|ref| {denom:.3e})"
    _close(out["pl_mean"].reshape(1), fx[f"{phase}/pl_mean_after"], "pl_mean", tol=1e-5)
    if "mapping.w_avg" in out["buffers"]:
        _close(out["buffers"]["mapping.w_avg"], fx[f"{phase}/w_avg_after"], "w_avg", tol=1e-5)
    else:
        assert np.array_equal(fx[f"{phase}/w_avg_after"], fx[f"{phase}/w_avg_before"])
