"""Inputs and sampling shared by the full-step fixture's generator (oracle/make_golden_r2.py, live reference) and its
checkers (tests/): deterministic pools for ``GAN_training_function.train`` and the strided sample stored per tensor."""
from __future__ import annotations

import torch

STEP_CFG = dict(resolution=32, G_ch=16, D_ch=16, G_attn="16", D_attn="16", n_classes=10, shared_dim=32,
                shared_dim_feat=64, class_cond=True, instance_cond=True)
# adam_eps is deliberately large: with 1e-6..1e-8 Adam maps a gradient that is pure rounding noise to a step of +-lr, two
# correct implementations then differ by O(lr) and the next step amplifies it (measured: 0.2 x lr after two steps).
STEP = dict(batch_size=4, n_acc=2, n_steps=3, G_lr=1e-3, D_lr=2e-3, B1=0.0, B2=0.999, adam_eps=1e-4, ema_decay=0.9,
            ema_start=2, seed=77)


def step_inputs(cfg, hp):
    """Per train() call a real batch of n_acc*batch_size rows; the conditioning sampler walks a pre-drawn pool
    (n_acc D draws + n_acc G draws per call)."""
    g = torch.Generator().manual_seed(hp["seed"])
    n = hp["batch_size"] * hp["n_acc"]
    calls = []
    for _ in range(hp["n_steps"]):
        x = torch.rand(n, 3, cfg.resolution, cfg.resolution, generator=g) * 2 - 1
        y = torch.randint(0, cfg.n_classes, (n,), generator=g)
        f = torch.nn.functional.normalize(torch.randn(n, cfg.feat_dim, generator=g), dim=1)
        calls.append((x, y, f))
    pool = []
    for _ in range(hp["n_steps"] * 2 * hp["n_acc"]):
        z = torch.randn(hp["batch_size"], cfg.eff_dim_z, generator=g)
        lab = torch.randint(0, cfg.n_classes, (hp["batch_size"],), generator=g)
        f = torch.nn.functional.normalize(torch.randn(hp["batch_size"], cfg.feat_dim, generator=g), dim=1)
        pool.append((z, lab, f))
    return calls, pool


def sample_of(t: torch.Tensor, limit=1024):
    """Full tensor if small, else a deterministic strided sample -- enough to pin a per-element update rule."""
    flat = t.detach().reshape(-1)
    if flat.numel() <= limit:
        return flat.clone()
    stride = flat.numel() // limit
    return flat[::stride][:limit].clone()
